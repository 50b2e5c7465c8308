"""The ospTestSuite image comparison metric, restated with numpy.

reference: apps/ospTestSuite/test_tools.h:26-29 (pixelThreshold 0.05, errorRate 0.0004, sumThreshold 9.0,
sumThresholdRatio 0.1) and test_tools.cpp:167-281 (OSPImageTools::getAveragedPixel /
compareImgWithBaselineTmpl): diff = baseline - rendered per RGBA channel; each pixel's error is the sum over
channels of |5x5 box average of diff| (window clipped at the borders) scaled by 1/255 for 8-bit images; it
counts only above pixelThreshold; meanError = sum / (4*W*H) must be <= errorRate. Unless the image is
denoised, the mean signed error must also stay <= sumThreshold and, when the mean |box-averaged| error
exceeds sumThreshold, the ratio of the two <= sumThresholdRatio."""
import numpy as np

PIXEL_THRESHOLD = 0.05
ERROR_RATE = 0.0004
SUM_THRESHOLD = 9.0
SUM_THRESHOLD_RATIO = 0.1


def _box5(img):
    """5x5 box average with the window clipped at the image borders (getAveragedPixel)"""
    h, w = img.shape[:2]
    pad = np.zeros((h + 5, w + 5) + img.shape[2:], np.float64)
    pad[3:h + 3, 3:w + 3] = img
    ii = pad.cumsum(axis=0).cumsum(axis=1)
    s = ii[5:, 5:] - ii[:-5, 5:] - ii[5:, :-5] + ii[:-5, :-5]
    ones = np.zeros((h + 5, w + 5), np.float64)
    ones[3:h + 3, 3:w + 3] = 1.0
    ci = ones.cumsum(axis=0).cumsum(axis=1)
    cnt = ci[5:, 5:] - ci[:-5, 5:] - ci[5:, :-5] + ci[:-5, :-5]
    return s / cnt.reshape(cnt.shape + (1,) * (img.ndim - 2))


def compare(rendered, baseline, denoised=False):
    """rendered/baseline: (H, W, 4) uint8 (PNG path, factor 1/255) or float32 (HDR path, factor 1).
    Returns a dict with `ok` and the statistics the reference prints."""
    rendered, baseline = np.asarray(rendered), np.asarray(baseline)
    assert rendered.shape == baseline.shape and rendered.shape[-1] == 4
    factor = 1.0 / 255.0 if rendered.dtype == np.uint8 else 1.0
    diff = baseline.astype(np.float32).astype(np.float64) - rendered.astype(np.float32).astype(np.float64)
    avg = np.abs(_box5(diff))
    pixel_error = avg.sum(axis=-1) * factor
    total = float(pixel_error[pixel_error > PIXEL_THRESHOLD].sum())
    rcp = 1.0 / (4.0 * rendered.shape[0] * rendered.shape[1])
    mean_error = total * rcp
    failed = mean_error > ERROR_RATE
    mean_sum = float(np.abs(diff.sum(axis=(0, 1))).sum()) * rcp * factor
    mean_sum_abs_avg = float(avg.sum()) * rcp * factor
    if not denoised:
        failed = failed or mean_sum > SUM_THRESHOLD
        if mean_sum_abs_avg > SUM_THRESHOLD:
            failed = failed or (mean_sum / mean_sum_abs_avg) > SUM_THRESHOLD_RATIO
    return {"ok": not failed, "mean_error": mean_error, "total_error": total, "mean_sum": mean_sum,
            "mean_sum_abs_avg": mean_sum_abs_avg, "pixel_perfect": bool(not diff.any())}


def rmse(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(np.mean((a - b) ** 2)))
