"""What peer-memory plumbing does this box offer?  (development probe, 2+ GPUs, one process per GPU)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P \
        tools/peer_probe.py

Checks, each in its own try block so that one failure does not hide the others:
  A. torch.distributed._symmetric_memory: rendezvous, peer buffer pointers, multicast pointer, a store into the
     peer's buffer, the handle's barrier, and the bandwidth of a copy kernel reading / writing peer memory;
  B. legacy cudaIpc handles (cudaMalloc + cudaIpcGetMemHandle / cudaIpcOpenMemHandle) through libcudart;
  C. stream memory operations (cuStreamWriteValue32 / cuStreamWaitValue32) on local and peer-mapped flags.
"""
import ctypes
import glob
import os
import sys
import time
import traceback

import torch
import torch.distributed as dist


def log(rank, *a):
    print(f"[rank {rank}]", *a, flush=True)


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    peer = (rank + 1) % world
    dist.barrier()

    # ---------------- A. torch symmetric memory
    hdl = None
    try:
        import torch.distributed._symmetric_memory as symm
        nel = 64 << 20  # 256 MB of float32
        t = symm.empty(nel, dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        log(rank, "symm_mem rendezvous ok: buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs], "signal_pad_ptrs",
            [hex(p) for p in hdl.signal_pad_ptrs], "multicast_ptr", hex(getattr(hdl, "multicast_ptr", 0) or 0),
            "buffer_size", hdl.buffer_size, "signal_pad_size", hdl.signal_pad_size)
        t.fill_(float(rank))
        hdl.barrier()
        pv = hdl.get_buffer(peer, (nel,), torch.float32)
        v = float(pv[12345])
        log(rank, "read of peer buffer:", v, "expected", float(peer))
        hdl.barrier()
        pv[:1024].fill_(100.0 + rank)  # store into the peer's memory
        hdl.barrier()
        log(rank, "value my peer wrote into my buffer:", float(t[5]), "expected", 100.0 + (rank - 1) % world)
        scratch = torch.empty(nel, dtype=torch.float32, device=dev)
        for name, fn in (("pull (local <- peer)", lambda: scratch.copy_(pv)), ("push (peer <- local)", lambda: pv.copy_(scratch))):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            hdl.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            log(rank, f"{name}: {nel * 4 / ms / 1e6:.1f} GB/s ({ms:.3f} ms for 256 MB, all ranks at once)")
            hdl.barrier()
        # latency of the handle's barrier vs an NCCL all-reduce of one float
        one = torch.zeros(1, device=dev)
        for name, fn in (("symm barrier", lambda: hdl.barrier()), ("nccl all_reduce(1 float)", lambda: dist.all_reduce(one))):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            log(rank, f"{name}: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us per call (back to back on the stream)")
    except Exception:  # noqa: BLE001
        log(rank, "symm_mem FAILED:\n" + traceback.format_exc())
    dist.barrier()

    # ---------------- B. legacy cudaIpc through libcudart
    try:
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*"))
        for sp in sys.path:
            cands += glob.glob(os.path.join(sp, "nvidia", "cuda_runtime", "lib", "libcudart.so*"))
        cands += glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        rt = ctypes.CDLL(cands[0])
        log(rank, "libcudart:", cands[0])
        ptr = ctypes.c_void_p()
        assert rt.cudaMalloc(ctypes.byref(ptr), ctypes.c_size_t(1 << 20)) == 0
        handle = (ctypes.c_ubyte * 64)()
        rc = rt.cudaIpcGetMemHandle(handle, ptr)
        log(rank, "cudaIpcGetMemHandle rc", rc)
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        rt.cudaMemset(ptr, ctypes.c_int(rank + 1), ctypes.c_size_t(1 << 20))
        torch.cuda.synchronize()
        dist.barrier()
        pptr = ctypes.c_void_p()

        class H(ctypes.Structure):
            _fields_ = [("b", ctypes.c_ubyte * 64)]

        h = H()
        ctypes.memmove(h.b, handles[peer], 64)
        rt.cudaIpcOpenMemHandle.argtypes = [ctypes.POINTER(ctypes.c_void_p), H, ctypes.c_uint]
        rc = rt.cudaIpcOpenMemHandle(ctypes.byref(pptr), h, 1)
        log(rank, "cudaIpcOpenMemHandle rc", rc, "peer ptr", hex(pptr.value or 0))
        if rc == 0:
            buf = (ctypes.c_ubyte * 16)()
            rc = rt.cudaMemcpy(buf, pptr, ctypes.c_size_t(16), ctypes.c_int(2))
            log(rank, "read through the IPC mapping rc", rc, "byte", buf[0], "expected", peer + 1)
            dist.barrier()
            rt.cudaIpcCloseMemHandle(pptr)
        dist.barrier()
        rt.cudaFree(ptr)
    except Exception:  # noqa: BLE001
        log(rank, "cudaIpc FAILED:\n" + traceback.format_exc())
    dist.barrier()

    # ---------------- C. stream memory operations on the symmetric buffer
    try:
        if hdl is None:
            raise RuntimeError("no symmetric buffer")
        try:
            from cuda.bindings import driver as cu
        except ImportError:
            from cuda import cuda as cu
        stream = torch.cuda.current_stream(dev).cuda_stream
        flags = hdl.get_buffer(rank, (16,), torch.int32)
        flags.zero_()
        torch.cuda.synchronize()
        hdl.barrier()
        my_flag_ptr = hdl.buffer_ptrs[rank]
        peer_flag_ptr = hdl.buffer_ptrs[peer]
        # write 7 into the peer's flag from my stream, then wait on my own flag until it reads >= 7
        (err,) = cu.cuStreamWriteValue32(stream, peer_flag_ptr, 7, 0)
        log(rank, "cuStreamWriteValue32 to peer:", err)
        (err,) = cu.cuStreamWaitValue32(stream, my_flag_ptr, 7, 1)  # 1 = CU_STREAM_WAIT_VALUE_GEQ
        log(rank, "cuStreamWaitValue32 on own flag:", err)
        t0 = time.time()
        torch.cuda.synchronize()
        log(rank, f"stream released after {1e3 * (time.time() - t0):.2f} ms; flag = {int(flags[0])}")
        # timing: ping-pong of write + wait, 50 rounds
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        hdl.barrier()
        e0.record()
        for i in range(50):
            cu.cuStreamWriteValue32(stream, peer_flag_ptr, 100 + i, 0)
            cu.cuStreamWaitValue32(stream, my_flag_ptr, 100 + i, 1)
        e1.record()
        torch.cuda.synchronize()
        log(rank, f"write+wait round: {e0.elapsed_time(e1) / 50 * 1e3:.1f} us")
    except Exception:  # noqa: BLE001
        log(rank, "stream mem ops FAILED:\n" + traceback.format_exc())
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("PEER_PROBE_DONE", flush=True)


if __name__ == "__main__":
    main()
