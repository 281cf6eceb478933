"""Host logic behind the shared likelihood rows (csrc/bnpc_tc_i8s.cuh): which chains of a wave
share a tile, through the C ABI (bnpc_ll_shared_plan) -- no device needed."""
import ctypes as C

import pytest


@pytest.fixture(scope='module')
def L():
    from bnpc_b200 import _lib
    _lib.build()
    return _lib.lib()


def plan(L, kpads, cells, words):
    n = len(kpads)
    I = C.c_int * 8
    group, col, ctas, slots = I(*[-1] * 8), I(*[-1] * 8), I(*[0] * 8), I(*[0] * 8)
    T, ng = C.c_int(0), C.c_int(0)
    L.ll_shared_plan(n, (C.c_int * n)(*kpads), cells, words, group, col, C.byref(T), ctas, slots, C.byref(ng))
    return list(group[:n]), list(col[:n]), T.value, list(ctas[:ng.value]), list(slots[:ng.value]), ng.value


def test_benchmark_wave_is_two_groups_of_four(L):
    # C3: 100k cells, 1000 mutations (W = 32 plane words), clusters padded to 24 / 32
    kp = [24, 32, 24, 32, 24, 32, 24, 32]
    group, col, T, ctas, slots, ng = plan(L, kp, 100000, 32)
    assert ng == 2 and group == [0, 0, 0, 0, 1, 1, 1, 1]
    # columns: two digits per padded cluster, chains side by side
    assert col == [0, 48, 112, 160, 0, 48, 112, 160]
    # 224 columns per group: one tile per supertile, one CTA per SM, tables streamed through the ring
    assert T == 1 and ctas == [148, 148]
    assert all(0 < s < 16 for s in slots) and all(s * 224 * 128 <= 200 * 1024 for s in slots)


def test_a_pair_with_small_tables_is_resident_with_two_tiles_per_chunk(L):
    group, col, T, ctas, slots, ng = plan(L, [24, 24], 100000, 32)
    assert ng == 1 and group == [0, 0] and col == [0, 48]
    assert T == 2                                   # 2 * 96 accumulator columns fit, 782 tiles >= 4 per SM
    assert slots == [16]                            # one slot per chunk of a row: loaded once
    assert ctas == [148]


def test_capacity_and_first_fit(L):
    # 64-cluster chains fill 128 columns each: two per group
    group, col, T, ctas, slots, ng = plan(L, [64, 64, 64], 100000, 32)
    assert group == [0, 0, 1] and col == [0, 128, 0] and T == 1
    # first fit: the small chain joins the first group that still has room
    group, col, T, _, _, ng = plan(L, [64, 56, 8, 8], 50000, 32)
    assert group == [0, 0, 0, 1] and col == [0, 128, 240, 0] and ng == 2
    # never more than 8 chains or 256 columns per group
    group, col, T, _, _, ng = plan(L, [8] * 8, 3000, 4)
    assert ng == 1 and col == [16 * i for i in range(8)]
    # few tiles: a CTA per tile, one tile per supertile even for a single chain
    group, col, T, ctas, slots, ng = plan(L, [24], 1000, 32)
    assert T == 1 and ctas == [8] and slots == [16]
    # long rows (C4: 5000 mutations, W = 160): the tables cannot stay resident
    _, _, T, ctas, slots, _ = plan(L, [24], 50000, 160)
    assert slots[0] < 80 and slots[0] <= 32


def test_bad_arguments(L):
    with pytest.raises(RuntimeError):
        plan(L, [20], 1000, 32)                     # padding not a multiple of 8
    with pytest.raises(RuntimeError):
        plan(L, [24], 1000, 30)                     # W not a multiple of 4
