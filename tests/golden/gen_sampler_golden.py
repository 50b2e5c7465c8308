#!/usr/bin/env python3
"""Generates tests/golden/samplers.json: known-answer vectors for the integer sampler sequences.

The reference ships NO numeric vectors for these, so they are produced here by a third,
independent restatement in pure Python integers of modules/cpu/math/random.ih (hashes :15-56, PCG32
:116-149, TEA :159-172, Owen hash :234-266, LDSampler :283-394), math/sobol.ih:38-80, math/halton.ih:38-86
and render/util.ih:33-60 -- written from the reference text, not from oracle.cpp or samplers.cuh. The Sobol'
matrices and Faure permutations come from tools/gen_tables.py (first principles; `--check` there compares
them with the reference's .inl tables). Values are stored as uint32 bit patterns / exact integers so that the
comparison is bit-exact; the final u32 -> float step is to_float_unorm = float32(u) * 2^-32.
Run:  python tests/golden/gen_sampler_golden.py   (needs nothing but this repo)"""
import json
import os
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
import gen_tables as GT  # noqa: E402

M = 0xFFFFFFFF
SOBOL = GT.sobol_rev_table()
P3, P5, P7 = GT.perm_table(3, 5), GT.perm_table(5, 3), GT.perm_table(7, 3)


def f32bits(x):
    return struct.unpack("<I", struct.pack("<f", x))[0]


def to_float_unorm_bits(u):
    # float32(u) (round to nearest even) * 2^-32 is exact after the conversion
    f = struct.unpack("<f", struct.pack("<f", float(u)))[0]
    return f32bits(f * 2.0 ** -32)


def rev(x):
    return int("{:032b}".format(x & M)[::-1], 2)


def murmur_mix(h, k):
    k = (k * 0xcc9e2d51) & M
    k = ((k << 15) | (k >> 17)) & M
    k = (k * 0x1b873593) & M
    h ^= k
    return ((((h << 13) | (h >> 19)) & M) * 5 + 0xe6546b64) & M


def murmur_fin(h):
    h ^= h >> 16
    h = (h * 0x85ebca6b) & M
    h ^= h >> 13
    h = (h * 0xc2b2ae35) & M
    h ^= h >> 16
    return h


def triple_hash(x):
    x ^= x >> 17
    x = (x * 0xed5ad4bb) & M
    x ^= x >> 11
    x = (x * 0xac4c1b51) & M
    x ^= x >> 15
    x = (x * 0x31848bab) & M
    x ^= x >> 14
    return x


def owen_hash2(x, seed):
    x ^= (x * 0x3d20adea) & M
    x = (x + seed) & M
    x = (x * ((seed >> 16) | 1)) & M
    x ^= (x * 0x05526c56) & M
    x ^= (x * 0x53a22864) & M
    return x


def owen_scramble4(x, seed):
    x = rev(x)
    x ^= (x * 0x3d20adea) & M
    x ^= (x >> 1) & ((x << 1) & M) & 0x55555555
    x = (x + seed) & M
    x = (x * ((seed >> 16) | 1)) & M
    x ^= (x >> 1) & ((x << 1) & M) & 0x55555555
    x ^= (x * 0x05526c56) & M
    x ^= (x * 0x53a22864) & M
    return rev(x)


def part8(n):
    n &= 0xFF
    n = (n | (n << 4)) & 0x0F0F0F0F
    n = (n | (n << 2)) & 0x33333333
    n = (n | (n << 1)) & 0x55555555
    return n


def sobol_rev(rev_index, dims):
    r = [0, 0, 0, 0]
    i = 0
    while rev_index:
        if rev_index & 0x80000000:
            for d in range(dims):
                r[d] ^= SOBOL[i + d]
        rev_index = (rev_index << 1) & M
        i += 4
    return r


S0, S1, S2, S3, S4 = 0x9374b0d3, 0x1ec38426, 0xdb81c29c, 0x5628372b, 0xbcb745bd


def scr(sample, scramble):
    return to_float_unorm_bits(rev(owen_hash2(sample, scramble)))


def ld_stream(seed, index):
    """the 12 values oracle_sampler_kat(kind 0) emits, as float bit patterns"""
    rev_index, scramble = rev(index), seed
    mask = (~(M >> 16)) & M
    scramble = triple_hash(scramble)                       # getNext5Samples: nextGroup
    ri = owen_hash2(rev_index, scramble) & mask
    r = sobol_rev(ri, 4)
    out = [scr(rev(ri), scramble ^ S0), scr(r[0], scramble ^ S1), scr(r[1], scramble ^ S2), scr(r[2], scramble ^ S3),
           scr(r[3], scramble ^ S4)]
    scramble = triple_hash(scramble)                       # nextGroup before NEE
    ri = owen_hash2(rev_index, scramble) & mask            # get3LightSamples(idx 0, split false)
    r = sobol_rev(ri, 2)
    out += [scr(rev(ri), scramble ^ S0), scr(r[0], scramble ^ S1), scr(r[1], scramble ^ S2)]
    scramble = triple_hash(scramble)                       # getNext4Samples: nextGroup
    ri = owen_hash2(rev_index, scramble) & mask
    r = sobol_rev(ri, 3)
    out += [scr(rev(ri), scramble ^ S0), scr(r[0], scramble ^ S1), scr(r[1], scramble ^ S2), scr(r[2], scramble ^ S3)]
    return out


def f32(x):
    return struct.unpack("<f", struct.pack("<f", x))[0]


def halton_stream(i):
    h2 = to_float_unorm_bits(rev(i))
    v3 = (P3[i % 243] * 14348907 + P3[(i // 243) % 243] * 59049 + P3[(i // 59049) % 243] * 243 + P3[(i // 14348907) % 243]) & M
    h3 = f32(f32(float(v3)) * f32(1.0 / 3486784401))
    x = f32(h3 + f32(1.0 / 6.0))                          # CranleyPattersonRotation(h3, 1/6)
    if x >= 1.0:
        x = f32(x - 1.0)
    v5 = (P5[i % 125] * 1953125 + P5[(i // 125) % 125] * 15625 + P5[(i // 15625) % 125] * 125 + P5[(i // 1953125) % 125]) & M
    v7 = (P7[i % 343] * 117649 + P7[(i // 343) % 343] * 343 + P7[(i // 117649) % 343]) & M
    return [h2, f32bits(x), f32bits(f32(f32(float(v5)) * f32(1.0 / 244140625))), f32bits(f32(f32(float(v7)) * f32(1.0 / 40353607)))]


def pcg_stream(seed, stream, n=4):
    M64 = (1 << 64) - 1
    state, st = 0, ((stream << 1) | 1) & M

    def step():
        nonlocal state
        old = state
        state = (old * 6364136223846793005 + (st | 1)) & M64
        xs = (((old >> 18) ^ old) >> 27) & M
        rot = old >> 59
        return ((xs >> rot) | (xs << ((-rot) & 31))) & M
    state = murmur_mix(state & M, seed)
    state = murmur_fin(state & M)
    step()
    state = (state + seed) & M64
    step()
    return [to_float_unorm_bits(step()) for _ in range(n)]


def tea8(v0, v1):
    s = 0
    for _ in range(8):
        s = (s + 0x9e3779b9) & M
        v0 = (v0 + ((((v1 << 4) & M) + 0xa341316c) & M ^ ((v1 + s) & M) ^ (((v1 >> 5) + 0xc8013ea4) & M))) & M
        v1 = (v1 + ((((v0 << 4) & M) + 0xad90777d) & M ^ ((v0 + s) & M) ^ (((v0 >> 5) + 0x7e95761e) & M))) & M
    return [to_float_unorm_bits(v0), to_float_unorm_bits(v1)]


def main():
    cases = {"ld": [], "halton": [], "pcg": [], "tea": [], "owen4": []}
    for seed in (0x49c5f37e, 0x49c5f37e ^ 12345, 1, 0xdeadbeef):
        for index in (0, 1, 2, 3, 7, 64, 1023, 65535, 65536, 1000003):
            cases["ld"].append({"seed": seed, "index": index, "bits": ld_stream(seed, index)})
    for i in (0, 1, 2, 3, 5, 17, 242, 243, 59049, 123456, 14348907, 4000000000):
        cases["halton"].append({"index": i, "bits": halton_stream(i)})
    for seed in (0, 1, 777, 0xffffffff):
        for stream in (0, 1, 5, 1000):
            cases["pcg"].append({"seed": seed, "stream": stream, "bits": pcg_stream(seed, stream)})
    for a in (0, 1, 100, 1023):
        for b in (0, 7, 767):
            cases["tea"].append({"v0": a, "v1": b, "bits": tea8(a, b)})
    for x in (0, 1, 5, 255):
        for y in (0, 3, 200):
            cases["owen4"].append({"x": x, "y": y, "value": owen_scramble4(part8(x) | (part8(y) << 1), 0x49c5f37e)})
    with open(os.path.join(HERE, "samplers.json"), "w") as f:
        json.dump(cases, f, indent=0)
    print("wrote samplers.json:", {k: len(v) for k, v in cases.items()})


if __name__ == "__main__":
    main()
