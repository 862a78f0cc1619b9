"""Thin runtime helpers over the C ABI: device selection, stream binding, sync."""
import ctypes as C

from ._lib import check, lib


def device_count():
    n = C.c_int()
    check(lib().smc_device_count(C.byref(n)))
    return n.value


def set_device(i):
    check(lib().smc_set_device(int(i)))


def set_stream(cuda_stream_ptr):
    """Bind this thread's launches to an existing cudaStream_t (e.g.
    ``torch.cuda.current_stream().cuda_stream``); 0/None -> library stream."""
    check(lib().smc_set_stream(C.c_void_p(cuda_stream_ptr or 0)))


def synchronize():
    check(lib().smc_synchronize())


def _point_at_bundled_nccl():
    """A Python process usually carries an NCCL of its own (the nvidia-nccl wheel
    torch links against).  The library must bind THAT copy, not the system's older
    one under the same soname -- whichever is loaded first is the one the whole
    process gets -- so SMC_NCCL_LIB names it before the first sharded call."""
    import os
    if os.environ.get("SMC_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            path = os.path.join(base, "lib", "libnccl.so.2")
            if os.path.exists(path):
                os.environ["SMC_NCCL_LIB"] = path
                return
    except Exception:
        pass


def shard_init(n_shards=0, devices=None):
    """Set up the shard set for row-sharded matrices: one shard per GPU (all visible
    ones by default); `devices` may repeat an id to put several shards on one GPU
    (host-side reduction then; for tests on a single-GPU box)."""
    _point_at_bundled_nccl()
    arr = None
    if devices is not None:
        n_shards = len(devices)
        arr = (C.c_int * n_shards)(*devices)
    check(lib().smc_shard_init(int(n_shards), arr))
    return shard_count()


def shard_count():
    n = C.c_int()
    check(lib().smc_shard_count(C.byref(n)))
    return n.value


def shard_reduce_mode():
    return (lib().smc_shard_reduce_mode() or b"").decode()


def shard_shutdown():
    check(lib().smc_shard_shutdown())


def timer_start():
    """CUDA events on the launching stream(s): device time of what follows."""
    check(lib().smc_timer_start())


def timer_stop():
    ms = C.c_double()
    check(lib().smc_timer_stop(C.byref(ms)))
    return ms.value


def measure_dmma_peak():
    """TFLOP/s of the FP64 tensor pipe from registers (a few milliseconds)."""
    v = C.c_double()
    check(lib().smc_measure_dmma_peak(C.byref(v)))
    return v.value


def device_info():
    sm, maj, mnr = C.c_int(), C.c_int(), C.c_int()
    fr, tot = C.c_size_t(), C.c_size_t()
    check(lib().smc_device_info(C.byref(sm), C.byref(maj), C.byref(mnr),
                                C.byref(fr), C.byref(tot)))
    return {"sm_count": sm.value, "cc": (maj.value, mnr.value),
            "free_bytes": fr.value, "total_bytes": tot.value}


def launch_count():
    return lib().smc_launch_count()


def reset_launch_count():
    lib().smc_reset_launch_count()
