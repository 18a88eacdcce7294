import torch
dev = torch.device("cuda", 0)
for n in (8192, 2048, 538):
    x = torch.rand((n, n), device=dev, dtype=torch.float64); y = torch.rand((n, n), device=dev, dtype=torch.float64)
    for _ in range(3):
        torch.matmul(x, y)
torch.cuda.synchronize()
