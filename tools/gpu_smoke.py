"""First-contact GPU script: render each test scene with each renderer, dump images + timings."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ospray_b200 import capi as osp, testing as T

def save_ppm(path, rgba):
    img = rgba[::-1, :, :3]
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img).tobytes())

def main():
    out = os.environ.get("OUT", "gpurun_out")
    os.makedirs(out, exist_ok=True)
    osp.init()
    W = H = int(os.environ.get("RES", "512"))
    res = {}
    scenes = {"cornell_box": T.cornell_box(), "boxes": T.boxes(), "random_spheres": T.random_spheres(),
              "gravity_spheres_volume": T.gravity_spheres_volume(64), "box_mesh": T.box_mesh(20)}
    for sname, spec in scenes.items():
        world, keep = T.to_ospray(spec)
        cam, pose = T.make_camera(world, W, H)
        print(sname, "bounds", world.bounds(), "cam", pose, flush=True)
        for rname in ("pathtracer", "scivis", "ao"):
            if sname == "gravity_spheres_volume" and rname == "pathtracer":
                continue
            ren = T.make_renderer(rname)
            fb = osp.FrameBuffer(W, H, osp.OSP_FB_SRGBA, osp.OSP_FB_COLOR | osp.OSP_FB_ACCUM | osp.OSP_FB_DEPTH | osp.OSP_FB_ID_PRIMITIVE)
            t0 = time.time()
            durs = []
            for i in range(4):
                f = fb.render_frame(ren, cam, world)
                f.wait()
                durs.append(f.duration() * 1e3)
            rc = fb.ray_counts()
            img = fb.map(osp.OSP_FB_COLOR)
            depth = fb.map(osp.OSP_FB_DEPTH)
            prim = fb.map(osp.OSP_FB_ID_PRIMITIVE)
            save_ppm(f"{out}/{sname}_{rname}.ppm", img)
            res[f"{sname}/{rname}"] = {"ms": durs, "rays": rc, "mean_rgb": img[..., :3].mean(axis=(0, 1)).tolist(),
                                       "depth_finite": float(np.isfinite(depth).mean()), "prims": int(len(np.unique(prim)))}
            print(sname, rname, res[f"{sname}/{rname}"], flush=True)
    json.dump(res, open(f"{out}/smoke.json", "w"), indent=1)

if __name__ == "__main__":
    main()
