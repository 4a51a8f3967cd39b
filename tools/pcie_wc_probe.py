"""Does write-combined pinned memory raise the H2D rate of this platform?  Duplex copies (H2D from a default / a
write-combined pinned buffer, D2H into a default pinned buffer) timed with CUDA events.  cuda-python's runtime bindings."""
from cuda import cudart


def ck(r):
    assert r[0] == cudart.cudaError_t.cudaSuccess, r[0]
    return r[1:] if len(r) > 2 else (r[1] if len(r) == 2 else None)


n = 1 << 30
ck(cudart.cudaSetDevice(0))
d_a, d_b = ck(cudart.cudaMalloc(n)), ck(cudart.cudaMalloc(n))
h_def = ck(cudart.cudaHostAlloc(n, cudart.cudaHostAllocDefault))
h_wc = ck(cudart.cudaHostAlloc(n, cudart.cudaHostAllocWriteCombined))
h_out = ck(cudart.cudaHostAlloc(n, cudart.cudaHostAllocDefault))
ck(cudart.cudaMemset(d_b, 1, n))
s1, s2 = ck(cudart.cudaStreamCreate()), ck(cudart.cudaStreamCreate())
e0, e1 = ck(cudart.cudaEventCreate()), ck(cudart.cudaEventCreate())
H2D, D2H = cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost


def run(src, duplex, reps=6):
    for timed in (False, True):
        ck(cudart.cudaDeviceSynchronize())
        ck(cudart.cudaEventRecord(e0, s1))
        for _ in range(reps if timed else 1):
            ck(cudart.cudaMemcpyAsync(d_a, src, n, H2D, s1))
            if duplex:
                ck(cudart.cudaMemcpyAsync(h_out, d_b, n, D2H, s2))
        ck(cudart.cudaStreamSynchronize(s2))
        ck(cudart.cudaEventRecord(e1, s1))
        ck(cudart.cudaDeviceSynchronize())
    return n * reps / (ck(cudart.cudaEventElapsedTime(e0, e1)) * 1e-3) / 1e9


for name, src in (("default pinned", h_def), ("write-combined", h_wc)):
    print(f"{name:15s} H2D alone {run(src, False):6.1f} GB/s   H2D while D2H runs {run(src, True):6.1f} GB/s per direction")
