"""Multi-GPU determinism check (run under torchrun, one rank per GPU):

  token-sharded calibration + reduce-to-owner  ==  single-process calibration   (<= 1e-6 relative)
  layer-distributed decompositions             ==  single-process decompositions (bit-exact files)

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
"""
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from modegpt_b200 import distributed as D  # noqa: E402
from modegpt_b200.adapters.CompressionConfig import CompressionConfig  # noqa: E402
from modegpt_b200.adapters.model_adapter import ModelAdapter  # noqa: E402
from modegpt_b200.calibration import load_calibs  # noqa: E402
from modegpt_b200.compression.compress_mlp import compress_nystrom  # noqa: E402
from modegpt_b200.compression.compress_qk import compress_qk  # noqa: E402
from modegpt_b200.compression.compress_vo import compress_vo  # noqa: E402
from modegpt_b200.compression_utils import allocate_global_sparsity  # noqa: E402
from modegpt_b200.eval import synthetic_tokens  # noqa: E402
from modegpt_b200.model_utils import build_synthetic_model  # noqa: E402


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def run(adapter, tokens, tag, tmp):
    L = adapter.n_layers
    adapter.calibs = [tokens[i:i + 2] for i in range(0, tokens.shape[0], 2)]
    adapter.config.temp_storage_dir = os.path.join(tmp, tag)
    adapter.config.keep_layers_in_memory = True
    adapter._layer_store = {}
    cov = load_calibs(adapter, tokens.shape[0], 2, dataset="synthetic", target_layers=list(range(L)))
    keep = allocate_global_sparsity(cov[4], 0.3, 0.04948, 0.95)
    compress_nystrom(adapter, cov[0], keep, list(range(L)))
    masks = compress_qk(adapter, (cov[1], cov[2]), keep, target_layers=list(range(L)))
    compress_vo(adapter, cov[3], keep, target_layers=list(range(L)))
    return cov, keep, masks, dict(adapter._layer_store)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
    rank, world = dist.get_rank(), dist.get_world_size()
    tmp = tempfile.mkdtemp(prefix="mg_dist_")
    model = build_synthetic_model("tiny-llama-gqa", device=dev, seed=0, n_layers=4, max_positions=512)
    adapter = ModelAdapter.from_model(model, None)
    adapter.config = CompressionConfig(model="tiny", order="mlp,qk,vo", ridge_vo=1e-5, ridge_qk=1e-2,
                                       nystrom_ridge=1e-4, dataset="synthetic", seq_len=256)
    tokens = synthetic_tokens(8, 256, model.config.vocab_size, 1234).to(dev)

    cov_d, keep_d, masks_d, layers_d = run(adapter, tokens, "dist", tmp)
    D._force_single = True
    cov_s, keep_s, masks_s, layers_s = run(adapter, tokens, "single", tmp)
    D._force_single = False

    worst = 0.0
    for l in range(adapter.n_layers):
        if l % world != rank:
            assert cov_d[0][l] is None, "non-owners must drop their statistics"
            continue
        for k in range(4):
            worst = max(worst, rel(cov_d[k][l], cov_s[k][l]))
        for suf in ("mlp", "qk", "vo"):
            for name, t in layers_d[(l, suf)].items():
                ref = layers_s[(l, suf)][name]
                assert t.shape == ref.shape
                if suf != "vo" and name != "down":
                    assert torch.equal(t, ref), (l, suf, name)     # gathers: bit-exact
                else:
                    assert rel(t, ref) < 1e-3, (l, suf, name, rel(t, ref))
    assert worst < 1e-5, worst
    assert max(abs(a - b) for a, b in zip(keep_d, keep_s)) < 1e-6
    assert len(masks_d) == adapter.n_layers
    for a, b in zip(masks_d, masks_s):
        assert torch.equal(a.cpu(), b.cpu())
    w = torch.tensor([worst], device=dev)
    dist.all_reduce(w, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"dist_check ok: world={world} worst statistic rel diff {w.item():.2e}; "
              f"keep ratios, masks and gathered weights identical")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
