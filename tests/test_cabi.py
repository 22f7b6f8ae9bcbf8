"""CPU-side checks of the C ABI's handle contract (include/poseengine.h): destroy calls are idempotent and ignore handles
the library does not know, pe_shutdown is callable any number of times -- no GPU needed."""
import ctypes as C

from posepipeline_b200 import _lib


def test_destroy_ignores_unknown_and_null_handles():
    lib = _lib.load()
    junk = (C.c_char * 4096)()                      # memory that was never a handle
    for fn in (lib.pe_engine_destroy, lib.pe_model_destroy, lib.pe_lifter_destroy):
        assert fn(None) == 0
        assert fn(C.cast(junk, C.c_void_p)) == 0    # not registered: must not be dereferenced
        assert fn(C.cast(junk, C.c_void_p)) == 0


def test_calls_on_dead_handles_fail_with_state_error():
    lib = _lib.load()
    junk = (C.c_char * 4096)()
    h = C.cast(junk, C.c_void_p)
    assert lib.pe_engine_sync(h) == _lib.PE_ERR_STATE
    n = C.c_int64()
    assert lib.pe_model_launch_count(h, C.byref(n)) == _lib.PE_ERR_STATE
    assert b"destroyed" in lib.pe_last_error()


def test_shutdown_is_idempotent():
    lib = _lib.load()
    assert lib.pe_shutdown() == 0 and lib.pe_shutdown() == 0
    _lib.shutdown()


def test_conv_schedule_visits_every_tile_and_slice_once():
    """conv_tc's persistent schedule (csrc/conv_tc_kernel.cuh tc_work_item, host view pe_tc_work_item): whatever the group size,
    the work items of a layer enumerate every (M tile, N slice) pair exactly once, a group's tiles finish all their slices before
    the next group starts, and a partial last group is handled.  Host-only: no GPU needed."""
    lib = _lib.load()
    tile, nsl = C.c_int32(), C.c_int32()

    def walk(cin, cout, ns, rows, tiles_m, mode, kb):
        seq, grp = [], None
        for w in range(tiles_m * ns):
            g = lib.pe_tc_work_item(cin, cout, ns, rows, tiles_m, mode, kb, w, C.byref(tile), C.byref(nsl))
            assert g >= 0 and (grp is None or g == grp)
            grp = g
            seq.append((tile.value, nsl.value))
        return grp, seq

    cases = [  # Cin, Cout, N slices, rows per tile, M tiles, mode, budget KB
        (768, 3072, 24, 256, 768, 1, 48 * 1024),     # ViT fc1: 393 KB of activations per tile pair -> groups of 64 tile pairs
        (3072, 768, 6, 512, 192, 1, 48 * 1024),      # ViT fc2 (CTA pairs, MT = 2)
        (96, 96, 2, 128, 75, 1, 700),                # small budget: partial last group
        (64, 256, 4, 128, 114, 2, 500),              # forced grouping of a layer the rule leaves n-major
        (1024, 1024, 8, 256, 37, 1, 1),              # budget below one tile: groups of one tile
        (48, 48, 1, 256, 100, 1, 48 * 1024),         # a single slice: n-major
    ]
    for cin, cout, ns, rows, tiles_m, mode, kb in cases:
        grp, seq = walk(cin, cout, ns, rows, tiles_m, mode, kb)
        assert len(set(seq)) == tiles_m * ns == len(seq), (cin, cout, ns)
        assert all(0 <= t < tiles_m and 0 <= n < ns for t, n in seq)
        if grp == 0:
            assert seq == [(w % tiles_m, w // tiles_m) for w in range(tiles_m * ns)]
        else:
            assert 1 <= grp <= tiles_m
            tile_kb = rows * (cin // 16) * (64 if lib.pe_precision_mode() == 1 else 128) / 1024
            assert grp == max(1, min(tiles_m, int(kb / tile_kb)))
            for w, (t, n) in enumerate(seq):                       # item w belongs to group w // (grp * ns), slices ascend inside it
                g = w // (grp * ns)
                assert g * grp <= t < min((g + 1) * grp, tiles_m)
            for g in range((tiles_m + grp - 1) // grp):
                part = seq[g * grp * ns:(g + 1) * grp * ns]
                assert [n for _, n in part] == sorted(n for _, n in part)
    # the rule of mode 1: grouped only where Cin * (ns - 1) >= Cout
    assert lib.pe_tc_work_item(64, 256, 2, 128, 114, 1, 48 * 1024, 0, C.byref(tile), C.byref(nsl)) == 0
    assert lib.pe_tc_work_item(64, 256, 2, 128, 114, 0, 48 * 1024, 0, C.byref(tile), C.byref(nsl)) == 0
    assert lib.pe_tc_work_item(96, 96, 2, 128, 75, 1, 48 * 1024, 0, C.byref(tile), C.byref(nsl)) == 75
    assert lib.pe_tc_work_item(96, 96, 2, 128, 75, 1, 48 * 1024, 150, C.byref(tile), C.byref(nsl)) < 0      # out of range
    assert lib.pe_tc_work_item(40, 96, 2, 128, 75, 1, 48 * 1024, 0, C.byref(tile), C.byref(nsl)) < 0        # Cin not a multiple of 16
