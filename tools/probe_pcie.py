"""What caps the host-buffer (e2e) leg when several ranks copy at once?  

    torchrun --nproc-per-node N tools/probe_pcie.py          (all N ranks copy concurrently, one GPU each)
    python tools/probe_pcie.py --one-process N               (ONE process drives N GPUs with one stream pair each)

Per rank and buffer flavour: H2D alone, D2H alone, both at once (GB/s each way), 512 MiB buffers, plain cudaMemcpyAsync.
  default   torch pin_memory=True (cudaHostAlloc default flags), first touched by this rank after CPU-affinity binding
  wc        cudaHostAllocWriteCombined for the H2D source (no CPU cache snooping on the device's reads)
  portable  cudaHostAllocPortable | cudaHostAllocMapped
"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

NBYTES = 512 << 20
REPS = 6


def cudart():
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            return ctypes.CDLL(name)
        except OSError:
            continue
    raise RuntimeError("libcudart not found")


def host_alloc(rt, nbytes, flags):
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc flags={flags}: error {rc}")
    return p


def copy_rates(rt, dev_idx, h_src, h_dst, d_a, d_b, barrier):
    """(h2d, d2h, duplex each way) GB/s with raw cudaMemcpyAsync on two streams."""
    torch.cuda.set_device(dev_idx)
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    H2D, D2H = 1, 2

    def run(h2d, d2h):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(REPS):
            if h2d:
                rt.cudaMemcpyAsync(ctypes.c_void_p(d_a), h_src, ctypes.c_size_t(NBYTES), H2D, ctypes.c_void_p(sa.cuda_stream))
            if d2h:
                rt.cudaMemcpyAsync(h_dst, ctypes.c_void_p(d_b), ctypes.c_size_t(NBYTES), D2H, ctypes.c_void_p(sb.cuda_stream))
        sa.synchronize()
        sb.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return REPS * NBYTES / dt / 1e9

    return run(True, False), run(False, True), run(True, True)


def bind(local_rank):
    try:
        import pynvml

        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
        h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        local = {i * 64 + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        use = local & os.sched_getaffinity(0)
        if use:
            os.sched_setaffinity(0, use)
        try:
            numa = pynvml.nvmlDeviceGetNumaNodeId(h)
        except Exception:
            numa = None
        return len(use), numa
    except Exception as exc:
        return None, repr(exc)


def main():
    rt = cudart()
    rt.cudaHostAlloc.restype = ctypes.c_int
    rt.cudaMemcpyAsync.restype = ctypes.c_int
    if "--one-process" in sys.argv:
        n = int(sys.argv[sys.argv.index("--one-process") + 1])
        import threading

        res = [None] * n
        bar = threading.Barrier(n)

        def worker(i):
            torch.cuda.set_device(i)
            d_a = torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{i}")
            d_b = torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{i}")
            hs, hd = host_alloc(rt, NBYTES, 0), host_alloc(rt, NBYTES, 0)
            ctypes.memset(hs, 1, NBYTES)
            ctypes.memset(hd, 1, NBYTES)
            res[i] = copy_rates(rt, i, hs, hd, d_a.data_ptr(), d_b.data_ptr(), bar.wait)

        ths = [threading.Thread(target=worker, args=(i,)) for i in range(n)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        print(f"one process, {n} GPUs, one thread + stream pair per GPU (default pinned): per-GPU GB/s (h2d, d2h, duplex each way)")
        for i, r in enumerate(res):
            print(f"  gpu{i}: {r[0]:6.1f} {r[1]:6.1f} {r[2]:6.1f}")
        print(f"  sum duplex each way: {sum(r[2] for r in res):.1f} GB/s")
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    torch.cuda.set_device(lr)
    ncpu, numa = bind(lr)

    def barrier():
        if dist is not None:
            dist.barrier()

    d_a = torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{lr}")
    d_b = torch.empty(NBYTES, dtype=torch.uint8, device=f"cuda:{lr}")
    rows = []
    for name, fs, fd in (("default", 0, 0), ("wc", 0x04, 0), ("portable", 0x01 | 0x02, 0x01 | 0x02)):
        hs, hd = host_alloc(rt, NBYTES, fs), host_alloc(rt, NBYTES, fd)
        ctypes.memset(hs, 1, NBYTES)
        ctypes.memset(hd, 1, NBYTES)
        rows.append((name,) + copy_rates(rt, lr, hs, hd, d_a.data_ptr(), d_b.data_ptr(), barrier))
        rt.cudaFreeHost(hs)
        rt.cudaFreeHost(hd)
    line = f"rank {rank}/{world} cpus={ncpu} numa={numa}: " + " | ".join(f"{n}: h2d {a:5.1f} d2h {b:5.1f} duplex {c:5.1f}" for n, a, b, c in rows)
    if dist is not None:
        out = [None] * world
        dist.all_gather_object(out, (line, [r[3] for r in rows]))
        if rank == 0:
            for l, _ in out:
                print(l)
            for k, n in enumerate(("default", "wc", "portable")):
                print(f"sum over ranks, duplex each way, {n}: {sum(o[1][k] for o in out):.1f} GB/s")
        dist.destroy_process_group()
    else:
        print(line)


if __name__ == "__main__":
    main()
