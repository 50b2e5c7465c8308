"""Procedural scenes of the reference's ospray_testing library, restated as neutral SceneSpec
objects plus adapters: `to_ospray()` replays a spec through the unchanged OSPRay C API (the same
call sequence the C++ builders issue), and tests feed the very same spec to the CPU oracle.

reference: apps/common/ospray_testing/builders/Builder.cpp:12-93 (world, transfer function),
CornellBox.cpp:26-251, GravitySpheresVolume.cpp:68-134,180-253, Boxes.cpp:40-105,
RandomSpheres.cpp:42-111; apps/common/arcball_camera/ArcballCamera.cpp:6-24,53-85;
apps/ospBenchmark/BaseFixture.cpp:18-50 (camera / framebuffer setup).
rkcommon's utility::uniform_real_distribution (un-vendored) is restated as
(g() - g.min()) * ((hi - lo) / float(g.max() - g.min())) + lo in float32."""
from dataclasses import dataclass, field
from typing import List, Optional
import numpy as np

from . import capi as osp


# ---- spec ---------------------------------------------------------------------------------------
@dataclass
class TextureSpec:
    data: np.ndarray                           # (h, w[, channels]) uint8 | uint16 | float16 | float32
    format: str                                # OSPTextureFormat name without the prefix: "RGBA8", "SRGB", "R32F", ...
    filter: str = "linear"                     # | "nearest"
    wrap: tuple = ("repeat", "repeat")         # repeat | mirror | clamp, per axis
    mipmaps: bool = True                       # False: rendered with the device parameter disableMipMapGeneration


@dataclass
class MapSpec:
    """a material's texture parameter with its optional 2D coordinate transform (render/Material.cpp:73-165)"""
    texture: TextureSpec
    translation: Optional[tuple] = None
    scale: Optional[tuple] = None
    rotation: Optional[float] = None           # degrees
    transform: Optional[tuple] = None          # linear2f (vx.x, vx.y, vy.x, vy.y)


@dataclass
class MaterialSpec:
    type: str = "obj"
    kd: tuple = (0.8, 0.8, 0.8)
    ks: tuple = (0.0, 0.0, 0.0)
    ns: float = 10.0
    d: float = 1.0
    tf: tuple = (0.0, 0.0, 0.0)
    eta: float = 1.5                           # thinGlass (render/materials/ThinGlass.cpp:33-48)
    attenuation_color: tuple = (1.0, 1.0, 1.0)
    attenuation_distance: float = 1.0
    thickness: float = 1.0
    map_kd: Optional["MapSpec"] = None
    map_ks: Optional["MapSpec"] = None
    map_ns: Optional["MapSpec"] = None
    map_d: Optional["MapSpec"] = None
    map_bump: Optional["MapSpec"] = None
    map_attenuation_color: Optional["MapSpec"] = None
    color: tuple = (1.0, 1.0, 1.0)              # luminous (render/materials/Luminous.cpp:33-45)
    intensity: float = 1.0
    transparency: float = 0.0
    map_color: Optional["MapSpec"] = None
    metal_eta: tuple = (1.47, 0.984, 0.602)     # metal (render/materials/Metal.cpp:30-76): complex index per channel (aluminium)
    metal_k: tuple = (7.64, 6.55, 5.36)
    roughness: float = 0.1                      # metal / plastic
    map_roughness: Optional["MapSpec"] = None
    pigment_color: tuple = (1.0, 1.0, 1.0)      # plastic (Plastic.cpp:28-38); `eta` above is its index of refraction
    eta_outside: float = 1.0                    # glass (Glass.cpp:35-63): `eta`, attenuation_color / _distance above = the inside medium
    attenuation_color_outside: tuple = (1.0, 1.0, 1.0)
    alloy_color: tuple = (0.9, 0.9, 0.9)        # alloy (Alloy.cpp:28-40): reflectivity at normal incidence / at grazing angles, `roughness`,
    edge_color: tuple = (1.0, 1.0, 1.0)         #   maps map_color / map_edge_color / map_roughness
    map_edge_color: Optional["MapSpec"] = None
    params: dict = field(default_factory=dict)   # velvet / metallicPaint / carPaint: ospray parameter name -> value (maps as MapSpec)
    principled: dict = field(default_factory=dict)  # principled (Principled.cpp:33-150): ospray parameter name -> value


@dataclass
class GeometrySpec:
    type: str                                  # mesh | sphere | box
    vertex: Optional[np.ndarray] = None        # (N,3) f32
    index: Optional[np.ndarray] = None         # (M,3|4) u32; None = soup
    quad_soup: bool = False
    vertex_color: Optional[np.ndarray] = None  # (N,4) f32
    vertex_normal: Optional[np.ndarray] = None
    vertex_texcoord: Optional[np.ndarray] = None  # (N,2) f32
    sphere_position: Optional[np.ndarray] = None
    sphere_radius: Optional[np.ndarray] = None
    radius: float = 0.01
    boxes: Optional[np.ndarray] = None         # (N,6) f32
    curve_vertex: Optional[np.ndarray] = None  # (N,4) f32 vertex.position_radius
    curve_index: Optional[np.ndarray] = None   # (M,) u32 first control point of each segment
    curve_basis: str = "catmull-rom"           # | "linear"


@dataclass
class ModelSpec:
    geometry: GeometrySpec
    material: Optional[MaterialSpec] = None
    color: Optional[np.ndarray] = None         # (P,4) per-primitive colour
    id: Optional[int] = None


@dataclass
class VolumeSpec:
    voxels: np.ndarray                         # (z,y,x) f32
    grid_origin: tuple = (0, 0, 0)
    grid_spacing: tuple = (1, 1, 1)
    tf_colors: Optional[np.ndarray] = None
    tf_opacities: Optional[np.ndarray] = None
    value_range: tuple = (0.0, 1.0)
    density_scale: float = 1.0
    gradient_shading_scale: float = 0.0
    anisotropy: float = 0.0                    # VolumetricModel 'anisotropy' (Henyey-Greenstein g, pathtracer only)


@dataclass
class GroupSpec:
    models: List[ModelSpec] = field(default_factory=list)
    volumes: List[VolumeSpec] = field(default_factory=list)
    lights: List["LightSpec"] = field(default_factory=list)   # group lights: they move with the instance


@dataclass
class InstanceSpec:
    group: GroupSpec
    xfm: Optional[np.ndarray] = None           # 12 floats column-major (vx, vy, vz, p)
    id: Optional[int] = None


@dataclass
class LightSpec:
    type: str
    params: dict = field(default_factory=dict)


@dataclass
class SceneSpec:
    instances: List[InstanceSpec] = field(default_factory=list)
    lights: List[LightSpec] = field(default_factory=list)
    backplate: Optional[TextureSpec] = None    # renderer parameter map_backplate
    max_depth_texture: Optional[TextureSpec] = None  # renderer parameter map_maxDepth (R32F, nearest)

    def uses_mipmaps(self):
        """False when any texture of the scene asks for disableMipMapGeneration (a device parameter: all or none)"""
        texs = [t for t in (self.backplate, self.max_depth_texture) if t is not None]
        texs += [v for ls in self.lights for v in ls.params.values() if isinstance(v, TextureSpec)]
        for inst in self.instances:
            for m in inst.group.models:
                if m.material is not None:
                    texs += [getattr(m.material, a).texture for a in ("map_kd", "map_ks", "map_ns", "map_d", "map_bump", "map_attenuation_color")
                             if getattr(m.material, a) is not None]
                    if m.material.map_color is not None: texs.append(m.material.map_color.texture)
                    if m.material.map_roughness is not None: texs.append(m.material.map_roughness.texture)
                    if m.material.map_edge_color is not None: texs.append(m.material.map_edge_color.texture)
                    texs += [v.texture for k, v in m.material.principled.items() if k.startswith("map_")]
                    texs += [v.texture for k, v in m.material.params.items() if k.startswith("map_")]
        return all(t.mipmaps for t in texs)


# ---- deterministic RNG ----------------------------------------------------------------------------
class Mt19937:
    """std::mt19937 stream (init_genrand seeding)"""

    def __init__(self, seed):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed))

    def raw(self, n):
        return self.bg.random_raw(n).astype(np.uint64)

    def uniform(self, lo, hi, n):
        lo, hi = np.float32(lo), np.float32(hi)
        scale = np.float32(hi - lo) / np.float32(4294967295.0)
        return (self.raw(n).astype(np.float32) * scale + lo).astype(np.float32)


# ---- builders ---------------------------------------------------------------------------------------
def _default_lights():
    # Builder::buildWorld: one invisible ambient light (Builder.cpp:43-47)
    return [LightSpec("ambient", {"visible": False})]


_CORNELL_QUADS = [
    # Floor, Ceiling, Backwall, RightWall, LeftWall
    [(1, -1, -1), (-1, -1, -1), (-1, -1, 1), (1, -1, 1)],
    [(1, 1, -1), (1, 1, 1), (-1, 1, 1), (-1, 1, -1)],
    [(1, -1, 1), (-1, -1, 1), (-1, 1, 1), (1, 1, 1)],
    [(-1, -1, 1), (-1, -1, -1), (-1, 1, -1), (-1, 1, 1)],
    [(1, -1, -1), (1, -1, 1), (1, 1, 1), (1, 1, -1)],
    # ShortBox: top, left, front, right, back, bottom
    [(-0.53, -0.40, -0.75), (-0.70, -0.40, -0.17), (-0.13, -0.40, -0.00), (0.05, -0.40, -0.57)],
    [(0.05, -1.00, -0.57), (0.05, -0.40, -0.57), (-0.13, -0.40, -0.00), (-0.13, -1.00, -0.00)],
    [(-0.53, -1.00, -0.75), (-0.53, -0.40, -0.75), (0.05, -0.40, -0.57), (0.05, -1.00, -0.57)],
    [(-0.70, -1.00, -0.17), (-0.70, -0.40, -0.17), (-0.53, -0.40, -0.75), (-0.53, -1.00, -0.75)],
    [(-0.13, -1.00, -0.00), (-0.13, -0.40, -0.00), (-0.70, -0.40, -0.17), (-0.70, -1.00, -0.17)],
    [(-0.53, -1.00, -0.75), (-0.70, -1.00, -0.17), (-0.13, -1.00, -0.00), (0.05, -1.00, -0.57)],
    # TallBox: top, left, front, right, back, bottom
    [(0.53, 0.20, -0.09), (-0.04, 0.20, 0.09), (0.14, 0.20, 0.67), (0.71, 0.20, 0.49)],
    [(0.53, -1.00, -0.09), (0.53, 0.20, -0.09), (0.71, 0.20, 0.49), (0.71, -1.00, 0.49)],
    [(0.71, -1.00, 0.49), (0.71, 0.20, 0.49), (0.14, 0.20, 0.67), (0.14, -1.00, 0.67)],
    [(0.14, -1.00, 0.67), (0.14, 0.20, 0.67), (-0.04, 0.20, 0.09), (-0.04, -1.00, 0.09)],
    [(-0.04, -1.00, 0.09), (-0.04, 0.20, 0.09), (0.53, 0.20, -0.09), (0.53, -1.00, -0.09)],
    [(0.53, -1.00, -0.09), (-0.04, -1.00, 0.09), (0.14, -1.00, 0.67), (0.71, -1.00, 0.49)],
]
_CORNELL_WHITE, _CORNELL_GREEN, _CORNELL_RED = (0.725, 0.710, 0.68, 1.0), (0.140, 0.450, 0.091, 1.0), (0.630, 0.065, 0.05, 1.0)


def cornell_box():
    """CornellBox.cpp:26-251: 17 quads as a quad soup with per-vertex colours, default obj material,
    one quad light that replaces the builder's ambient light."""
    verts = np.array([v for q in _CORNELL_QUADS for v in q], dtype=np.float32)
    cols = []
    for qi in range(len(_CORNELL_QUADS)):
        c = _CORNELL_GREEN if qi == 3 else _CORNELL_RED if qi == 4 else _CORNELL_WHITE
        cols += [c] * 4
    geom = GeometrySpec("mesh", vertex=verts, quad_soup=True, vertex_color=np.array(cols, dtype=np.float32))
    group = GroupSpec(models=[ModelSpec(geom, MaterialSpec())])
    light = LightSpec("quad", {"color": (0.78, 0.551, 0.183), "intensity": 47.0, "position": (-0.23, 0.98, -0.16),
                               "edge1": (0.47, 0.0, 0.0), "edge2": (0.0, 0.0, 0.38)})
    return SceneSpec(instances=[InstanceSpec(group)], lights=[light])


def boxes(dimensions=(4, 4, 4), lit=False):
    """Boxes.cpp:40-105: a lattice of box primitives with per-primitive colours."""
    dim = max(dimensions)
    bl, cl = [], []
    # index_sequence_3D iterates x fastest
    for k in range(dimensions[2]):
        for j in range(dimensions[1]):
            for i in range(dimensions[0]):
                f = np.array([i, j, k], dtype=np.float32)
                lower = f * np.float32(5.0)
                upper = lower + np.float32(0.75 * 5.0)
                bl.append(np.concatenate([lower, upper]))
                c = (np.float32(0.8) * f / np.float32(dim)) + np.float32(0.2)
                cl.append([c[0], c[1], c[2], 1.0])
    geom = GeometrySpec("box", boxes=np.array(bl, dtype=np.float32))
    model = ModelSpec(geom, MaterialSpec(ks=(0.3, 0.3, 0.3), ns=10.0), color=np.array(cl, dtype=np.float32))
    lights = _default_lights()
    if lit:
        lights = [LightSpec("distant", {"color": (0.78, 0.551, 0.483), "intensity": 3.14, "direction": (-0.8, -0.6, 0.3)}),
                  LightSpec("ambient", {"intensity": 0.35, "visible": False})]
    return SceneSpec(instances=[InstanceSpec(GroupSpec(models=[model]))], lights=lights)


def _ground_plane(bounds_lower, bounds_upper):
    """Builder::makeGroundPlane (Builder.cpp:95-200): a quad plus 2x10 stripes at y = -1."""
    lo, hi = np.asarray(bounds_lower, np.float32), np.asarray(bounds_upper, np.float32)
    center = 0.5 * (lo + hi)
    extent = np.float32(0.8) * np.float32(np.linalg.norm(center - lo))
    pos, col, idx = [], [], []
    plane_color, stripe_color = (0.9, 0.9, 0.9, 1.0), (1.0, 0.1, 0.1, 1.0)

    def quad(p0, p1, p2, p3, c):
        s = len(pos)
        pos.extend([p0, p1, p2, p3])
        col.extend([c] * 4)
        idx.append([s, s + 1, s + 2, s + 3])
    quad((-extent, -1, -extent), (extent, -1, -extent), (extent, -1, extent), (-extent, -1, extent), plane_color)
    sw = np.float32(0.025)
    padded = extent + sw
    y = np.float32(-1.0 + 1e-3)
    for i in range(10):
        coord = -extent + np.float32(i) / np.float32(9) * np.float32(2.0) * extent
        quad((coord - sw, y, -padded), (coord + sw, y, -padded), (coord + sw, y, padded), (coord - sw, y, padded), stripe_color)
        quad((-padded, y, coord - sw), (padded, y, coord - sw), (padded, y, coord + sw), (-padded, y, coord + sw), stripe_color)
    n = len(pos)
    geom = GeometrySpec("mesh", vertex=np.array(pos, np.float32), index=np.array(idx, np.uint32),
                        vertex_color=np.array(col, np.float32), vertex_normal=np.tile(np.array([[0, 1, 0]], np.float32), (n, 1)))
    return InstanceSpec(GroupSpec(models=[ModelSpec(geom, MaterialSpec())]))


def random_spheres(num_spheres=100, seed=0, add_plane=True, renderer_type="scivis"):
    """RandomSpheres.cpp:42-111: per-sphere colours; default obj material, or thinGlass with attenuationDistance 0.2 when
    built for the pathtracer (:89-99)."""
    gen = Mt19937(seed)
    centers = gen.uniform(-1, 1, 3 * num_spheres).reshape(num_spheres, 3)
    radii = gen.uniform(0.05, 0.15, num_spheres)
    colors = gen.uniform(0.5, 1.0, 4 * num_spheres).reshape(num_spheres, 4)
    geom = GeometrySpec("sphere", sphere_position=centers, sphere_radius=radii)
    mat = MaterialSpec(type="thinGlass", attenuation_distance=0.2) if renderer_type == "pathtracer" else MaterialSpec()
    group = GroupSpec(models=[ModelSpec(geom, mat, color=colors)])
    inst = [InstanceSpec(group)]
    if add_plane:
        lo = (centers - radii[:, None]).min(axis=0)
        hi = (centers + radii[:, None]).max(axis=0)
        inst.append(_ground_plane(lo, hi))
    return SceneSpec(instances=inst, lights=_default_lights())


def streamlines(seed=0, num_lines=100):
    """Streamlines.cpp:31-128: spirals of round Catmull-Rom curves with tapering radius, one colour per line, obj material.
    The builder's arithmetic is replayed in its types (int degrees, double sin/cos arguments, float results)."""
    f = np.float32
    gen = Mt19937(seed)
    # the reference draws every number from ONE generator, alternating between five distributions: one raw value per draw
    def draw(lo, hi):
        return gen.uniform(lo, hi, 1)[0]
    points, indices, colors = [], [], []

    def lerp(t, a, b):
        return (f(1.0) - f(t)) * a + f(t) * b

    for _ in range(num_lines):
        d_start, d_end = int(draw(0, 360)), int(draw(360, 720))
        radius, h, h_step, fr = draw(0.5, 1.5), f(0.0), draw(0.001, 0.1), draw(0.5, 1.5)
        r = f(720 - d_end) / f(360.0)
        c = np.array([r, f(1) - r, f(1) - r / f(2), f(1.0)], np.float32)
        d = d_start
        while d < d_start + d_end:
            p = np.array([radius * np.sin(d * np.pi / f(180.0)), h - f(2), radius * np.cos(d * np.pi / f(180.0))]).astype(np.float32)
            start_radius = f(f(0.015) * np.sin(np.float64(fr) * d * np.pi / 180) + f(0.02))
            q = np.array([(radius - f(0.05)) * np.sin((d + 10) * np.pi / f(180.0)), h + h_step - f(2),
                          (radius - f(0.05)) * np.cos((d + 10) * np.pi / f(180.0))]).astype(np.float32)
            end_radius = f(f(0.015) * np.sin(np.float64(fr) * (d + 10) * np.pi / 180) + f(0.02))
            if d == d_start:
                rim = lerp(f(1.0) + end_radius / f(np.linalg.norm(q - p)), q, p)
                cap = lerp(f(1.0) + start_radius / f(np.linalg.norm(rim - p)), p, rim)
                points += [np.append(cap, f(0)), np.append(rim, f(0)), np.append(p, start_radius), np.append(q, end_radius)]
                indices.append(len(points) - 4)
                colors += [c, c]
            elif d + 10 < d_start + d_end and d + 20 > d_start + d_end:
                rim = lerp(f(1.0) + start_radius / f(np.linalg.norm(p - q)), p, q)
                cap = lerp(f(1.0) + end_radius / f(np.linalg.norm(rim - q)), q, rim)
                points += [np.append(p, start_radius), np.append(q, end_radius), np.append(rim, f(0)), np.append(cap, f(0))]
                indices += [len(points) - 7, len(points) - 6, len(points) - 5, len(points) - 4]
                colors += [c, c]
            elif (d != d_start and d != d_start + 10) and d + 20 < d_start + d_end:
                points.append(np.append(p, start_radius))
                indices.append(len(points) - 4)
            colors.append(c)
            radius = radius - f(0.05)
            d += 10
            h = h + h_step
    geom = GeometrySpec("curve", curve_vertex=np.array(points, np.float32), curve_index=np.array(indices, np.uint32),
                        vertex_color=np.array(colors, np.float32)[:len(points)], curve_basis="catmull-rom")
    return SceneSpec(instances=[InstanceSpec(GroupSpec(models=[ModelSpec(geom, MaterialSpec())]))], lights=_default_lights())


def jet_transfer_function():
    colors = np.array([(0, 0, 0.562493), (0, 0, 1), (0, 1, 1), (0.500008, 1, 0.500008), (1, 1, 0), (1, 0, 0),
                       (0.500008, 0, 0)], dtype=np.float32)
    opacities = np.array([0.0, 1.0], dtype=np.float32)
    return colors, opacities


def gravity_spheres_voxels(dim=128, num_points=10, seed=0):
    """GravitySpheresVolume.cpp:180-232"""
    gen = Mt19937(seed)
    pts = np.zeros((num_points, 4), dtype=np.float32)
    for p in range(num_points):  # interleaved draws: x, y, z (centre dist) then weight
        pts[p, 0:3] = gen.uniform(-1, 1, 3)
        pts[p, 3] = gen.uniform(0.1, 0.3, 1)[0]
    ax = (np.float32(-1.0) + np.arange(dim, dtype=np.float32) / np.float32(dim - 1) * np.float32(2.0)).astype(np.float32)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    vox = np.zeros((dim, dim, dim), dtype=np.float32)
    for p in range(num_points):
        dx, dy, dz = x - pts[p, 0], y - pts[p, 1], z - pts[p, 2]
        d2 = (dx * dx + dy * dy + dz * dz).astype(np.float32)
        dist = np.sqrt(d2).astype(np.float32)
        with np.errstate(divide="ignore"):
            vox += (pts[p, 3] / (dist * dist)).astype(np.float32)
    return vox


def gravity_spheres_volume(dim=128, seed=0):
    """GravitySpheresVolume.cpp:68-134,234-253: one structuredRegular volume, jet/linear TF over [0,10]."""
    vox = gravity_spheres_voxels(dim, 10, seed)
    colors, opac = jet_transfer_function()
    s = np.float32(2.0) / np.float32(dim)
    vol = VolumeSpec(vox, grid_origin=(-1.0, -1.0, -1.0), grid_spacing=(float(s),) * 3, tf_colors=colors, tf_opacities=opac,
                     value_range=(0.0, 10.0))
    return SceneSpec(instances=[InstanceSpec(GroupSpec(volumes=[vol]))], lights=_default_lights())


def box_mesh(n_boxes_per_axis, seed=0, jitter=0.25):
    """synthetic stress mesh: K^3 jittered boxes, 12 triangles each, one indexed mesh."""
    K = int(n_boxes_per_axis)
    gen = Mt19937(seed)
    nb = K * K * K
    ii, jj, kk = np.meshgrid(np.arange(K), np.arange(K), np.arange(K), indexing="ij")
    base = np.stack([ii.ravel(), jj.ravel(), kk.ravel()], axis=1).astype(np.float32)
    off = gen.uniform(-jitter, jitter, 3 * nb).reshape(nb, 3)
    size = gen.uniform(0.25, 0.45, nb).astype(np.float32)
    c = (base + off + 0.5) / np.float32(K) * 2.0 - 1.0
    h = (size / np.float32(K))[:, None]
    corners = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], np.float32)
    verts = (c[:, None, :] + corners[None, :, :] * h[:, None, :]).reshape(-1, 3).astype(np.float32)
    tri = np.array([[0, 2, 1], [0, 3, 2], [4, 5, 6], [4, 6, 7], [0, 1, 5], [0, 5, 4], [2, 3, 7], [2, 7, 6], [1, 2, 6], [1, 6, 5],
                    [0, 4, 7], [0, 7, 3]], np.uint32)
    index = (tri[None, :, :] + (np.arange(nb, dtype=np.uint32) * 8)[:, None, None]).reshape(-1, 3).astype(np.uint32)
    colors = np.concatenate([0.3 + 0.7 * (c * 0.5 + 0.5), np.ones((nb, 1), np.float32)], axis=1).astype(np.float32)
    vcol = np.repeat(colors, 8, axis=0)
    geom = GeometrySpec("mesh", vertex=verts, index=index, vertex_color=vcol)
    return SceneSpec(instances=[InstanceSpec(GroupSpec(models=[ModelSpec(geom, MaterialSpec())]))], lights=_default_lights())


BUILDERS = {"cornell_box": cornell_box, "boxes": boxes, "random_spheres": random_spheres,
            "gravity_spheres_volume": gravity_spheres_volume, "box_mesh": box_mesh}


# ---- arcball camera ------------------------------------------------------------------------------------
def arcball_camera(bounds_lower, bounds_upper):
    """ArcballCamera ctor + eyePos/lookDir/upDir for the un-rotated camera (ArcballCamera.cpp:6-24,53-71)"""
    lo, hi = np.asarray(bounds_lower, np.float32), np.asarray(bounds_upper, np.float32)
    diag = (hi - lo).astype(np.float32)
    diag = np.maximum(diag, np.float32(0.3) * np.float32(np.linalg.norm(diag))).astype(np.float32)
    center = (np.float32(0.5) * (lo + hi)).astype(np.float32)
    dist = np.float32(np.linalg.norm(diag))
    eye = center + np.array([0, 0, np.float32(1.0) - dist], dtype=np.float32)
    return {"position": tuple(float(v) for v in eye), "direction": (0.0, 0.0, 1.0), "up": (0.0, 1.0, 0.0)}


# ---- replay a spec through the OSPRay API -----------------------------------------------------------------
def _geometry(g: GeometrySpec):
    geo = osp.Geometry(g.type)
    if g.type == "mesh":
        geo.set("vertex.position", g.vertex)
        if g.vertex_color is not None:
            geo.set("vertex.color", g.vertex_color)
        if g.vertex_normal is not None:
            geo.set("vertex.normal", g.vertex_normal)
        if g.vertex_texcoord is not None:
            geo.set("vertex.texcoord", np.ascontiguousarray(g.vertex_texcoord, np.float32))
        if g.index is not None:
            geo.set("index", g.index)
        elif g.quad_soup:
            geo.set("quadSoup", True)
    elif g.type == "sphere":
        geo.set("sphere.position", g.sphere_position)
        if g.sphere_radius is not None:
            geo.set("sphere.radius", g.sphere_radius)
        else:
            geo.set("radius", float(g.radius))
    elif g.type == "box":
        geo.set("box", g.boxes)
    elif g.type == "curve":
        geo.set("vertex.position_radius", np.ascontiguousarray(g.curve_vertex, np.float32))
        geo.set("index", np.ascontiguousarray(g.curve_index, np.uint32))
        if g.vertex_color is not None:
            geo.set("vertex.color", g.vertex_color)
        geo.set("type", osp.Typed(osp.OSP_UINT, 0))                                               # OSP_ROUND
        geo.set("basis", osp.Typed(osp.OSP_UINT, {"linear": 0, "bezier": 1, "bspline": 2, "catmull-rom": 4}[g.curve_basis]))  # OSPCurveBasis
    return geo.commit()


_FILTER = {"linear": 0, "nearest": 1}
_WRAP = {"repeat": 0, "mirror": 1, "clamp": 2}


def make_texture(t: TextureSpec):
    return osp.texture2d(t.data, t.format, _FILTER[t.filter], tuple(_WRAP[w] for w in t.wrap))


def _set_map(mat, name, ms: "MapSpec"):
    if ms is None:
        return
    mat.set(name, make_texture(ms.texture))
    if ms.translation is not None: mat.set(name + ".translation", tuple(float(x) for x in ms.translation))
    if ms.scale is not None: mat.set(name + ".scale", tuple(float(x) for x in ms.scale))
    if ms.rotation is not None: mat.set(name + ".rotation", float(ms.rotation))
    if ms.transform is not None: mat.set(name + ".transform", osp.Typed(osp.OSP_LINEAR2F, ms.transform))


def _material(m: MaterialSpec):
    mat = osp.Material(m.type)
    d = MaterialSpec()
    for name in ("map_kd", "map_ks", "map_ns", "map_d", "map_bump"):
        _set_map(mat, name, getattr(m, name))
    _set_map(mat, "map_attenuationColor", m.map_attenuation_color)
    if m.type == "metal":
        _set_map(mat, "map_roughness", m.map_roughness)
        mat.set("eta", tuple(float(x) for x in m.metal_eta)).set("k", tuple(float(x) for x in m.metal_k)).set("roughness", float(m.roughness))
        return mat.commit()
    if m.type == "plastic":
        mat.set("pigmentColor", tuple(float(x) for x in m.pigment_color)).set("eta", float(m.eta)).set("roughness", float(m.roughness))
        return mat.commit()
    if m.type == "glass":
        mat.set("eta", float(m.eta)).set("attenuationColor", tuple(float(x) for x in m.attenuation_color))
        mat.set("attenuationDistance", float(m.attenuation_distance))
        if m.eta_outside != d.eta_outside: mat.set("etaOutside", float(m.eta_outside))
        if m.attenuation_color_outside != d.attenuation_color_outside: mat.set("attenuationColorOutside", tuple(float(x) for x in m.attenuation_color_outside))
        return mat.commit()
    if m.type == "alloy":
        _set_map(mat, "map_color", m.map_color)
        _set_map(mat, "map_edgeColor", m.map_edge_color)
        _set_map(mat, "map_roughness", m.map_roughness)
        mat.set("color", tuple(float(x) for x in m.alloy_color)).set("edgeColor", tuple(float(x) for x in m.edge_color)).set("roughness", float(m.roughness))
        return mat.commit()
    if m.type in ("velvet", "metallicPaint", "carPaint"):
        for name, v in m.params.items():
            if name.startswith("map_"): _set_map(mat, name, v)
            elif isinstance(v, (tuple, list)): mat.set(name, tuple(float(x) for x in v))
            else: mat.set(name, float(v))
        return mat.commit()
    if m.type == "principled":
        for name, v in m.principled.items():
            if name.startswith("map_"): _set_map(mat, name, v)
            elif isinstance(v, bool): mat.set(name, bool(v))
            elif isinstance(v, (tuple, list)): mat.set(name, tuple(float(x) for x in v))
            else: mat.set(name, float(v))
        return mat.commit()
    if m.type == "luminous":
        _set_map(mat, "map_color", m.map_color)
        mat.set("color", tuple(float(x) for x in m.color))
        if m.intensity != d.intensity: mat.set("intensity", float(m.intensity))
        if m.transparency != d.transparency: mat.set("transparency", float(m.transparency))
        return mat.commit()
    if m.type == "thinGlass":
        if m.eta != d.eta: mat.set("eta", float(m.eta))
        if m.attenuation_color != d.attenuation_color: mat.set("attenuationColor", tuple(float(x) for x in m.attenuation_color))
        if m.attenuation_distance != d.attenuation_distance: mat.set("attenuationDistance", float(m.attenuation_distance))
        if m.thickness != d.thickness: mat.set("thickness", float(m.thickness))
        return mat.commit()
    # a parameter with a map defaults to factor 1 (Material.cpp:110-133): the spec's values are always sent then
    if m.kd != d.kd or m.map_kd: mat.set("kd", tuple(float(x) for x in m.kd))
    if m.ks != d.ks or m.map_ks: mat.set("ks", tuple(float(x) for x in m.ks))
    if m.ns != d.ns or m.ks != d.ks or m.map_ns: mat.set("ns", float(m.ns))
    if m.d != d.d: mat.set("d", float(m.d))
    if m.tf != d.tf: mat.set("tf", tuple(float(x) for x in m.tf))
    return mat.commit()


def to_ospray(scene: SceneSpec):
    """returns (world, keepalive list)"""
    keep = []
    instances = []
    for inst in scene.instances:
        models, vmodels = [], []
        for ms in inst.group.models:
            geo = _geometry(ms.geometry)
            model = osp.GeometricModel(geo)
            if ms.material is not None:
                model.set("material", _material(ms.material))
            if ms.color is not None:
                model.set("color", ms.color)
            if ms.id is not None:
                model.set("id", osp.Typed(osp.OSP_UINT, ms.id))
            models.append(model.commit())
        for vs in inst.group.volumes:
            vol = osp.Volume("structuredRegular")
            vol.set("gridOrigin", tuple(float(x) for x in vs.grid_origin))
            vol.set("gridSpacing", tuple(float(x) for x in vs.grid_spacing))
            vol.set("data", vs.voxels)
            vol.commit()
            tf = osp.TransferFunction("piecewiseLinear")
            tf.set("color", vs.tf_colors).set("opacity", vs.tf_opacities)
            tf.set("value", osp.Typed(osp.OSP_BOX1F, vs.value_range)).commit()
            vm = osp.VolumetricModel(vol)
            vm.set("transferFunction", tf)
            if vs.density_scale != 1.0: vm.set("densityScale", float(vs.density_scale))
            if vs.gradient_shading_scale != 0.0: vm.set("gradientShadingScale", float(vs.gradient_shading_scale))
            if vs.anisotropy != 0.0: vm.set("anisotropy", float(vs.anisotropy))
            vmodels.append(vm.commit())
            keep += [vol, tf]
        group = osp.Group()
        if models: group.set("geometry", models)
        if vmodels: group.set("volume", vmodels)
        glights = []
        for ls in inst.group.lights:
            l = osp.Light(ls.type)
            for k, v in ls.params.items():
                l.set(k, make_texture(v) if isinstance(v, TextureSpec) else v)
            glights.append(l.commit())
        if glights: group.set("light", glights)
        keep += glights
        group.commit()
        oi = osp.Instance(group)
        if inst.xfm is not None:
            oi.set("transform", osp.Typed(osp.OSP_AFFINE3F, np.asarray(inst.xfm, np.float32)))
        if inst.id is not None:
            oi.set("id", osp.Typed(osp.OSP_UINT, inst.id))
        instances.append(oi.commit())
        keep += models + vmodels + [group]
    world = osp.World()
    if instances:
        world.set("instance", instances)
    lights = []
    for ls in scene.lights:
        l = osp.Light(ls.type)
        for k, v in ls.params.items():
            l.set(k, make_texture(v) if isinstance(v, TextureSpec) else v)
        lights.append(l.commit())
    if lights:
        world.set("light", lights)
    world.commit()
    keep += instances + lights
    return world, keep


def make_camera(world, width, height, kind="perspective", ortho_height=1.0, stereo_mode=0, interpupillary_distance=None):
    """BaseFixture::Init camera (BaseFixture.cpp:36-47)"""
    lo, hi = world.bounds()
    p = arcball_camera(lo, hi)
    cam = osp.Camera(kind)
    if kind == "orthographic":
        cam.set("height", float(ortho_height))
    if stereo_mode:
        cam.set("stereoMode", osp.Typed(osp.OSP_UINT, int(stereo_mode)))
    if interpupillary_distance is not None:
        cam.set("interpupillaryDistance", float(interpupillary_distance))
    if kind != "panoramic":  # (the panoramic camera has no aspect: a full sphere whatever the image shape)
        cam.set("aspect", float(np.float32(width) / np.float32(height)))
    cam.set("position", p["position"]).set("direction", p["direction"]).set("up", p["up"])
    return cam.commit(), p


def make_renderer(kind, **params):
    r = osp.Renderer(kind)
    for k, v in params.items():
        r.set(k, v)
    return r.commit()
