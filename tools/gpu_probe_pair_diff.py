import os, sys, subprocess
import torch
sys.path.insert(0, ".")
if len(sys.argv) > 1:
    from modegpt_b200 import ops
    torch.manual_seed(0)
    out = {}
    for (T, n) in [(4096, 1024), (1000, 512), (8192, 2048), (4096, 768)]:
        X = (torch.randn(T, n, device="cuda") * torch.exp(0.5 * torch.randn(n, device="cuda"))).bfloat16()
        C = torch.zeros(n, n, device="cuda")
        ops.syrk_(C, X); ops.syrk_(C, X)
        ref = 2 * (X.double().T @ X.double())
        out[(T, n)] = (C.cpu(), ref.cpu())
    torch.save(out, sys.argv[1])
else:
    env = dict(os.environ)
    subprocess.check_call([sys.executable, __file__, "/tmp/pair.pt"], env=env)
    env["MG_DISABLE_2CTA"] = "1"
    subprocess.check_call([sys.executable, __file__, "/tmp/single.pt"], env=env)
    a, b = torch.load("/tmp/pair.pt"), torch.load("/tmp/single.pt")
    for k in a:
        cp, ref = a[k]; cs, _ = b[k]
        n = cp.shape[0]
        up = torch.triu(torch.ones(n, n, dtype=torch.bool))
        ep = ((cp.double() - ref)[up].norm() / ref[up].norm()).item()
        es = ((cs.double() - ref)[up].norm() / ref[up].norm()).item()
        d = (cp.double() - cs.double()) * up
        rel_tile = []
        for i in range(0, n, 256):
            row = []
            for j in range(0, n, 256):
                blk = d[i:i+256, j:j+256]; r = ref[i:i+256, j:j+256]
                row.append(f"{(blk.norm() / r.norm()).item():8.1e}")
            rel_tile.append(" ".join(row))
        print(f"T,n={k}: pair err {ep:.2e} single err {es:.2e}; per-256-tile |pair-single|/|ref|:")
        print("\n".join("    " + r for r in rel_tile))
