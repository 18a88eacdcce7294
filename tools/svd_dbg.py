"""In-kernel timers of the one-CTA Jacobi SVD (cycles / ns inside the sweep loop, sweeps) from a DEBUG build:
    nvcc -DT10_SVD_DEBUG -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --cudart shared \
         -Iinclude -Itor10_b200/csrc -shared tor10_b200/csrc/svd_small.cu tor10_b200/csrc/blockmap.cu \
         -o tools/bin/libsvddbg.so          (tools/bin is git-ignored; it travels to the GPU box with gpurun)
    python tools/svd_dbg.py [library name under tools/bin]"""
import ctypes as C, torch, json, sys, os
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "bin", sys.argv[1] if len(sys.argv) > 1 else "libsvddbg.so"))
dev = torch.device("cuda", 0)
for n in (64, 32, 16):
    x = torch.rand((n, n), device=dev, dtype=torch.float64)
    u = torch.empty((n, n), device=dev, dtype=torch.float64); s = torch.empty(n, device=dev, dtype=torch.float64); vh = torch.empty((n, n), device=dev, dtype=torch.float64)
    info = torch.zeros(8, device=dev, dtype=torch.int32)
    prob = (C.c_int64 * 9)(n, n, 0, n, 1, 0, 0, 0, -1)
    for _ in range(3):
        rc = lib.t10_svd_small(0, C.c_int64(1), prob, C.c_void_p(x.data_ptr()), C.c_void_p(u.data_ptr()), C.c_void_p(s.data_ptr()), C.c_void_p(vh.data_ptr()), None, C.c_void_p(info.data_ptr()), None)
    torch.cuda.synchronize()
    v = info.cpu().tolist()
    sweeps = v[0] >> 1
    rounds = sweeps * (n - 1)
    print(json.dumps({"n": n, "rc": rc, "sweeps": sweeps, "cycles": v[1], "ns": v[2], "cycles_per_round": v[1] / rounds, "ns_per_round": v[2] / rounds, "sm_clock_ghz": v[1] / max(v[2], 1)}))
