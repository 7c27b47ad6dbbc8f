"""Where does the layer-file hand-off cost time?  The three decomposition stages over 16
Llama-2-7B-shaped layers: in memory, and with files under several writer configurations
(MG_WRITER_SAVERS / MG_WRITER_STAGERS) and type-I worker counts."""
import os, shutil, sys, tempfile, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
from modegpt_b200.adapters.CompressionConfig import CompressionConfig
from modegpt_b200.adapters.model_adapter import ModelAdapter
from modegpt_b200.compression.compress_mlp import compress_nystrom
from modegpt_b200.compression.compress_qk import compress_qk
from modegpt_b200.compression.compress_vo import compress_vo
from modegpt_b200.model_utils import build_synthetic_model

sys.setswitchinterval(5e-4)
dev = "cuda:0"
L = 16
model = build_synthetic_model("llama-2-7b", device=dev, n_layers=L)
torch.manual_seed(0)
n, d, H, hd, T = 11008, 4096, 32, 128, 16384
x = (torch.randn(T, n, device=dev) * torch.exp(0.5 * torch.randn(n, device=dev))).bfloat16()
c = torch.zeros(n, n, device=dev); ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / T)
cx = torch.zeros(d, d, device=dev); ops.syrk_(cx, x[:, :d].contiguous()); ops.finalize_sym_(cx, 1.0 / T)
ch = torch.zeros(H, hd, hd, device=dev); ops.syrk_heads_(ch, x[:, :d].contiguous()); ops.scale_(ch, 1.0 / T)
del x
cov_mlp, cov_x, cov_q, cov_k, keep = [c] * L, [cx] * L, [ch] * L, [ch] * L, [0.75] * L
layers = list(range(L))


def run(files, workers, savers=8, stagers=3):
    os.environ["MG_WRITER_SAVERS"], os.environ["MG_WRITER_STAGERS"] = str(savers), str(stagers)
    adapter = ModelAdapter.from_model(model, None)
    tmp = tempfile.mkdtemp(prefix="mg_probe_layers_")
    adapter.config = CompressionConfig(model="x", order="mlp,qk,vo", nystrom_ridge=1e-4, ridge_vo=1e-5,
                                       ridge_qk=1e-2, keep_layers_in_memory=not files, mlp_workers=workers,
                                       temp_storage_dir=tmp)
    adapter.prepare_writer()
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for name, fn in (("mlp", lambda: compress_nystrom(adapter, cov_mlp, keep, layers)),
                     ("qk", lambda: compress_qk(adapter, (cov_q, cov_k), keep, target_layers=layers)),
                     ("vo", lambda: compress_vo(adapter, cov_x, keep, target_layers=layers))):
        t1 = time.perf_counter(); fn(); torch.cuda.synchronize(); t[name] = 1e3 * (time.perf_counter() - t1) / L
    t1 = time.perf_counter(); adapter.flush_saves(); t["flush_total_ms"] = 1e3 * (time.perf_counter() - t1)
    total = 1e3 * (time.perf_counter() - t0) / L
    if adapter._writer is not None:
        adapter._writer.close()
    adapter._layer_cache.clear(); adapter._layer_store.clear()
    shutil.rmtree(tmp, ignore_errors=True)
    print(f"files={files!s:5} workers={workers} savers={savers:2d} stagers={stagers}: {total:6.2f} ms/layer  "
          + " ".join(f"{k}={v:.2f}" for k, v in t.items()), flush=True)


def run_concurrent(vo_group):
    """mlp (2 workers) on the calling thread; qk + vo on a second thread / stream."""
    import threading
    adapter = ModelAdapter.from_model(model, None)
    adapter.config = CompressionConfig(model="x", order="mlp,qk,vo", nystrom_ridge=1e-4, ridge_vo=1e-5,
                                       ridge_qk=1e-2, keep_layers_in_memory=True, mlp_workers=2, vo_group=vo_group)
    err = []
    side = torch.cuda.Stream()
    def attn():
        try:
            torch.cuda.set_device(dev)
            with torch.no_grad(), torch.cuda.stream(side):
                compress_qk(adapter, (cov_q, cov_k), keep, target_layers=layers)
                compress_vo(adapter, cov_x, keep, target_layers=layers)
            side.synchronize()
        except BaseException as e:
            err.append(e)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    th = threading.Thread(target=attn); th.start()
    compress_nystrom(adapter, cov_mlp, keep, layers)
    t_mlp = time.perf_counter() - t0
    th.join(); torch.cuda.synchronize()
    total = 1e3 * (time.perf_counter() - t0) / L
    assert not err, err
    print(f"concurrent stages (vo_group={vo_group}): {total:6.2f} ms/layer  (mlp alone finished at {1e3 * t_mlp / L:.2f})", flush=True)
    adapter._layer_store.clear()


run(False, 2)
run(False, 2)
if os.environ.get("CONCURRENT"):
    for g in (1, 2, 4, 1, 2):
        run_concurrent(g)
cfgs = os.environ.get("CFGS", "2,8,3;2,4,2;2,2,1;2,12,3;1,8,3;1,2,1")
for cfg in cfgs.split(";"):
    run(True, *(int(v) for v in cfg.split(",")))
run(False, 1)
