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
