"""Oracle vs the committed known-answer vectors (tests/golden/samplers.json, produced by the pure-Python
restatement tests/golden/gen_sampler_golden.py). Bit-exact: integer hashing, then one u32 -> float step."""
import json
import os

import numpy as np

from oracle import pyoracle as O

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "samplers.json")))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_ldsampler_stream_bit_exact():
    # LDSampler_getNext5Samples/finalize, nextGroup, get3LightSamples, getNext4Samples (random.ih:283-394)
    for c in GOLD["ld"]:
        got = bits(O.sampler_kat(0, c["seed"], c["index"], 1)[0])
        assert got.tolist() == c["bits"], c


def test_halton_bit_exact():
    # Halton_sample2/3, PermHalton_sample5/7 (halton.ih:49-86) + CranleyPattersonRotation (random.ih:84-91)
    for c in GOLD["halton"]:
        got = bits(O.sampler_kat(1, 0, c["index"], 1)[0])
        assert got.tolist() == c["bits"], c


def test_pcg32_bit_exact():
    # RandomSampler_init / pcg32 (random.ih:116-149)
    for c in GOLD["pcg"]:
        got = bits(O.sampler_kat(2, c["seed"], c["stream"], 1)[0])
        assert got.tolist() == c["bits"], c


def test_tea8_bit_exact():
    for c in GOLD["tea"]:
        got = bits(O.sampler_kat(3, c["v0"], c["v1"], 1)[0])
        assert got.tolist() == c["bits"], c


def test_owen_scramble4_zorder():
    # OwenScramble4(interleaveZOrder16(x, y), seed) (random.ih:252-266, render/util.ih:57-60)
    for c in GOLD["owen4"]:
        got = int(O.sampler_kat(4, c["y"], c["x"], 1)[0, 0])
        assert got == c["value"], c


def test_sampler_ranges_and_stratification():
    # properties: values in [0,1]; the first two LD dimensions of 2^k consecutive indices are a (0,2)-net
    s = O.sampler_kat(0, 0x49c5f37e ^ 4242, 0, 256)
    assert (s >= 0).all() and (s <= 1).all()
    for k in (16, 64, 256):
        cell = np.floor(s[:k, 0] * k).astype(int)
        assert len(np.unique(cell)) == k  # dimension 0 is a radical inverse: one sample per 1/k stratum
    h = O.sampler_kat(1, 0, 1, 243)
    assert len(np.unique(np.floor(h[:, 1] * 243 * (1 - 1e-7)).astype(int))) == 243  # Halton base 3: full period
