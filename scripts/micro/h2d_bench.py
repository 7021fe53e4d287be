"""H2D bandwidth of one packed 90 MB step input: torch pinned memory vs write-combined pinned memory
(cudaHostAlloc with cudaHostAllocWriteCombined), and split into 2 / 4 chunks on separate streams."""
import ctypes
import torch

n = 90431040
rt = ctypes.CDLL("libcudart.so.12")
dev = torch.empty(n, dtype=torch.uint8, device="cuda")


def bench(host_ptr, label, chunks=1):
    streams = [torch.cuda.Stream() for _ in range(chunks)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sz = n // chunks
    for rep in range(3):
        torch.cuda.synchronize()
        e0.record()
        for it in range(10):
            for c, s in enumerate(streams):
                s.wait_stream(torch.cuda.current_stream())
                rc = rt.cudaMemcpyAsync(ctypes.c_void_p(dev.data_ptr() + c * sz), ctypes.c_void_p(host_ptr + c * sz),
                                        ctypes.c_size_t(sz), 1, ctypes.c_void_p(s.cuda_stream))
                assert rc == 0
            for s in streams:
                torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("%-40s %.3f ms  %.1f GB/s" % (label, ms, n / ms / 1e6))


pinned = torch.empty(n, dtype=torch.uint8).pin_memory()
pinned.fill_(1)
bench(pinned.data_ptr(), "torch pinned")
bench(pinned.data_ptr(), "torch pinned, 2 chunks / 2 streams", 2)
bench(pinned.data_ptr(), "torch pinned, 4 chunks / 4 streams", 4)
p = ctypes.c_void_p()
for flags, name in ((0x04, "write-combined"), (0x01, "portable"), (0x00, "default cudaHostAlloc")):
    assert rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), flags) == 0
    ctypes.memset(p.value, 1, n)
    bench(p.value, "cudaHostAlloc %s" % name)
    rt.cudaFreeHost(p)
