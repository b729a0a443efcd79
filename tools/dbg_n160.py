import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import mambo_b200 as mb
dev = torch.device("cuda", 0)
for n in (153, 160, 168):
    g = torch.Generator(device=dev).manual_seed(n)
    G = torch.randn((n, n, n, n), generator=g, dtype=torch.float64, device=dev)
    G = G + G.permute(1, 0, 2, 3)
    G = G + G.permute(0, 1, 3, 2)
    G = (G + G.permute(2, 3, 0, 1)).mul_(1.0 / (n * n)).contiguous()
    print(n, "torch asym", float((G - G.permute(2, 3, 0, 1)).abs().max()), float((G - G.permute(1, 0, 2, 3)).abs().max()), G.is_contiguous())
    try:
        with mb.CSAEngine.from_device(G.data_ptr(), n) as e:
            print(n, e.info())
    except Exception as ex:
        print(n, "ERR", ex)
    del G
    torch.cuda.empty_cache()
