#!/usr/bin/env python
"""Host<->device copy ceiling of this box: bare bidirectional pinned cudaMemcpyAsync, per GPU set and per buffer kind.

    python tools/pcie_ceiling.py [--seconds 0.6] [--mb 118]

For GPU sets {0}, {0,1}, {0,4}, {0,1,2,3}, {0,4,1,5}, {0..7} (as far as the box has them), one host thread per GPU
(bound to that GPU's NUMA-local CPUs before it allocates) streams `mb` MB chunks H2D and D2H on two CUDA streams at
once.  Buffer kinds: "pinned" (cudaHostAlloc), "huge" (2 MB transparent huge pages + cudaHostRegister), "wc"
(cudaHostAllocWriteCombined for the H2D side).  Prints one JSON line per (set, kind): GB/s per direction per GPU and in
aggregate — the denominators for bench.py's end-to-end numbers (118 MB = one 32-frame 720p YUYV batch)."""
import argparse
import ctypes
import json
import mmap
import os
import threading
import time

import torch


def cudart():
    """libcudart through ctypes (torch's own binding does not export cudaMemcpyAsync / cudaHostRegister)"""
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            L = ctypes.CDLL(name)
            L.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
            L.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
            L.cudaHostUnregister.argtypes = [ctypes.c_void_p]
            return L
        except OSError:
            continue
    import glob
    for path in glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib", "libcudart.so*")):
        L = ctypes.CDLL(path)
        L.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
        L.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
        L.cudaHostUnregister.argtypes = [ctypes.c_void_p]
        return L
    raise RuntimeError("libcudart not found")


RT = None


def numa_cpus(dev):
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        n = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n)
        return {i * 64 + b for i, m in enumerate(words) for b in range(64) if (m >> b) & 1}
    except Exception:
        return set()


class HostBuf:
    def __init__(self, nbytes, kind, write_combined=False):
        self.kind, self.nbytes = kind, nbytes
        rt = RT
        if kind == "huge":
            size = (nbytes + (2 << 20) - 1) // (2 << 20) * (2 << 20)
            self.mm = mmap.mmap(-1, size, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            try:
                self.mm.madvise(mmap.MADV_HUGEPAGE)
            except Exception:
                pass
            buf = (ctypes.c_char * size).from_buffer(self.mm)
            self.ptr = ctypes.addressof(buf)
            ctypes.memset(self.ptr, 1, size)                       # first touch on this thread's NUMA node
            r = rt.cudaHostRegister(self.ptr, size, 0)
            assert int(r) == 0, f"cudaHostRegister -> {r}"
            self.t = None
            self.size = size
        else:
            self.t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            self.t.fill_(1)
            self.ptr = self.t.data_ptr()


def run_set(gpus, kind, mb, seconds):
    nbytes = mb << 20
    res, errs = {}, []

    def worker(dev):
        try:
            cpus = numa_cpus(dev) & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
            torch.cuda.set_device(dev)
            hin, hout = HostBuf(nbytes, kind), HostBuf(nbytes, kind)
            din = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
            dout = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
            s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            rt = RT

            def burst(n):
                for _ in range(n):
                    rt.cudaMemcpyAsync(din.data_ptr(), hin.ptr, nbytes, 1, s1.cuda_stream)
                    rt.cudaMemcpyAsync(hout.ptr, dout.data_ptr(), nbytes, 2, s2.cuda_stream)
            burst(2); torch.cuda.synchronize(dev)
            barrier.wait()
            t0 = time.perf_counter(); n = 0
            while time.perf_counter() - t0 < seconds:
                burst(4); n += 4
                s1.synchronize(); s2.synchronize()
            dt = time.perf_counter() - t0
            res[dev] = n * nbytes / dt / 1e9
            for hb in (hin, hout):
                if hb.kind == "huge":
                    rt.cudaHostUnregister(hb.ptr)
        except Exception as e:                                         # report, never hang the barrier
            errs.append(f"gpu {dev}: {e}")
            try:
                barrier.abort()
            except Exception:
                pass
    barrier = threading.Barrier(len(gpus))
    th = [threading.Thread(target=worker, args=(g,)) for g in gpus]
    for t in th: t.start()
    for t in th: t.join()
    if errs:
        return {"gpus": gpus, "kind": kind, "error": errs[:2]}
    return {"gpus": gpus, "kind": kind, "gbs_each_direction_per_gpu": [round(res[g], 1) for g in gpus],
            "gbs_each_direction_total": round(sum(res.values()), 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=0.6)
    ap.add_argument("--mb", type=int, default=118)
    a = ap.parse_args()
    global RT
    torch.cuda.init()
    RT = cudart()
    n = torch.cuda.device_count()
    sets = [[0]]
    if n >= 2: sets.append([0, 1])
    if n >= 8: sets += [[0, 4], [0, 1, 2, 3], [0, 4, 1, 5], list(range(8))]
    elif n >= 4: sets.append([0, 1, 2, 3])
    full = os.sched_getaffinity(0)
    for s in sets:
        for kind in ("pinned", "huge"):
            os.sched_setaffinity(0, full)
            print(json.dumps(run_set(s, kind, a.mb, a.seconds)), flush=True)


if __name__ == "__main__":
    main()
