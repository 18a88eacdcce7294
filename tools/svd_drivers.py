import time, torch
dev = torch.device("cuda", 0)
for n in (64, 256, 1024):
    M = torch.rand((n, n), device=dev, dtype=torch.float64)
    for drv in (None, "gesvd", "gesvdj", "gesvda"):
        try:
            for _ in range(3):
                torch.linalg.svd(M, full_matrices=False, driver=drv)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            reps = 20 if n <= 256 else 3
            for _ in range(reps):
                u, s, vh = torch.linalg.svd(M, full_matrices=False, driver=drv)
            torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
            err = float(((u * s) @ vh - M).abs().max())
            print(n, drv, "%.3f ms" % (dt * 1e3), "recon err %.1e" % err)
        except Exception as e:
            print(n, drv, "failed", str(e)[:80])
