"""CPU: Philox4x32-10 known-answer vectors (Random123 kat_vectors) for the two statements of the generator that the
checks rely on: tests/philox.py (numpy; the C5 point stream of hpsdf_uniform_points_device is compared with it on the GPU)
and the C restatement's (oracle/hp_oracle.c; the sample points of the mc_counter nearness estimator)."""
import ctypes as C

import numpy as np

import philox

KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_numpy_philox_known_answers():
    for ctr, key, want in KAT:
        got = philox.philox4x32_10(np.array([ctr], np.uint32), np.array(key, np.uint32))[0]
        assert tuple(int(x) for x in got) == want


def test_oracle_philox_known_answers(oracle):
    L = oracle.lib()
    L.hporacle_philox.argtypes = [C.POINTER(C.c_uint32), C.c_uint32, C.c_uint32]
    for ctr, key, want in KAT:
        c = (C.c_uint32 * 4)(*ctr)
        L.hporacle_philox(c, key[0], key[1])
        assert tuple(c) == want
