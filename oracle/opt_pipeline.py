"""Oracle-level OPT pipeline (BASELINE config #1) — TEST INFRASTRUCTURE ONLY.

The reference's OPT path is dead in HEAD (its OPTAdapter cannot be
instantiated), so there is no reference run to record.  What survives is the intended semantics,
piece by piece; this module composes the pinned oracle functions accordingly, on CPU:

  statistics   C_mlp from relu(fc1(x)) == the input of fc2   (src/adapters/model_adapter.py:546-554)
               per-head C_q / C_k from the raw q_proj / k_proj outputs, biases included  (:556-567)
               C_x from the attention input (output of self_attn_layer_norm; the role
               on_batch_end_step was meant to play, src/adapters/OPTAdapter.py:45-46)
               Block-Influence as in src/calibration.py:118-136 (last pair ends at the decoder's
               final_layer_norm output, which is what HF returns as hidden_states[L])
  allocation   allocate_global_sparsity                          (src/compression_utils.py:79-124)
  type I       compress_weights without a gate; fc1 bias follows its rows, fc2 bias kept
               (src/compression/compress_mlp.py:28-64, src/adapters/model_adapter.py:442-452)
  type II      compress_head_opt per head, q/k bias entries follow their rows
               (src/compression/compress_qk.py:439-476)
  type III     compress_head (MHA, two SVDs) per head; out_proj bias kept
               (src/compression/compress_vo.py:162-223, src/adapters/model_adapter.py:529-538).
               The reference drops the V bias (bias=False for V, :529); attention weights sum to
               one, so the V bias reaches the output as the constant W_o b_v — this build folds it
               into the kept output bias instead (o_bias = W_o b_v + b_o): "parity unpinned" for
               that one vector, stated here and in DESIGN.md.

`run(model, batches, hyper)` takes an HF OPTForCausalLM on CPU and returns every intermediate and
every layer tensor, so a GPU run can be compared stage by stage.
"""
from __future__ import annotations

import numpy as np
import torch

from . import modegpt_oracle as O


@torch.no_grad()
def run(model, batches: list[torch.Tensor], compression_ratio: float, nystrom_ridge: float,
        ridge_vo: float, smoothing: float, max_sparsity: float) -> dict:
    cfg = model.config
    L, H, d = cfg.num_hidden_layers, cfg.num_attention_heads, cfg.hidden_size
    hd = d // H
    dec = model.model.decoder
    # statistics are accumulated batch by batch inside the hooks (raw sums, normalised below);
    # only the current batch's block inputs / outputs are kept, for the Block-Influence pairs
    acc = {"mlp": [0.0] * L, "x": [0.0] * L, "q": [0.0] * L, "k": [0.0] * L}
    cur: dict = {}
    handles = []
    f64 = lambda t: t.detach().to(torch.float64).numpy()

    def add(key, i, value):
        acc[key][i] = acc[key][i] + value

    for i, blk in enumerate(dec.layers):
        handles.append(blk.fc2.register_forward_pre_hook(
            lambda m, inp, i=i: add("mlp", i, O.gram_rows(f64(inp[0])))))
        handles.append(blk.self_attn_layer_norm.register_forward_hook(
            lambda m, inp, out, i=i: add("x", i, O.gram_rows(f64(out)))))
        handles.append(blk.self_attn.q_proj.register_forward_hook(
            lambda m, inp, out, i=i: add("q", i, O.gram_heads(f64(out), H, hd))))
        handles.append(blk.self_attn.k_proj.register_forward_hook(
            lambda m, inp, out, i=i: add("k", i, O.gram_heads(f64(out), H, hd))))
        handles.append(blk.register_forward_pre_hook(
            lambda m, args, kwargs, i=i: cur.__setitem__(("in", i), f64(args[0] if args else kwargs["hidden_states"])),
            with_kwargs=True))
        handles.append(blk.register_forward_hook(
            lambda m, args, out, i=i: cur.__setitem__(("out", i), f64(out[0] if isinstance(out, (tuple, list)) else out))))
    handles.append(dec.final_layer_norm.register_forward_hook(lambda m, inp, out: cur.__setitem__("final", f64(out))))
    bi = np.zeros(L)
    try:
        for b in batches:
            cur.clear()
            model.model(b, use_cache=False)
            B = b.shape[0]
            for l in range(L):
                x_out = cur["final"] if l == L - 1 else cur[("out", l)]
                bi[l] += O.bi_batch(cur[("in", l)].reshape(B, -1, d), x_out.reshape(B, -1, d))
    finally:
        for h in handles:
            h.remove()
    cur.clear()

    n_texts = sum(len(b) for b in batches)
    out: dict = {"n_texts": n_texts}
    bi /= n_texts
    keep = O.allocate_global_sparsity(bi, compression_ratio, smoothing, max_sparsity)
    out["bi"], out["keep"] = bi, np.array(keep)
    sd = {k: v.detach().float().numpy() for k, v in model.state_dict().items()}
    for l in range(L):
        pre = f"model.decoder.layers.{l}."
        c_mlp = O.normalise_stats(acc["mlp"][l], n_texts)
        c_x = O.normalise_stats(acc["x"][l], n_texts)
        c_q = O.normalise_stats(acc["q"][l], n_texts)
        c_k = O.normalise_stats(acc["k"][l], n_texts)
        out[f"cov_mlp{l}"], out[f"cov_x{l}"], out[f"cov_q{l}"], out[f"cov_k{l}"] = c_mlp, c_x, c_q, c_k
        # type I
        mlp, idx, rank = O.nystrom_mlp(sd[pre + "fc1.weight"], None, sd[pre + "fc2.weight"], c_mlp, keep[l],
                                       nystrom_ridge)
        out[f"L{l}_mlp_up"], out[f"L{l}_mlp_down"], out[f"L{l}_mlp_idx"] = mlp["up"], mlp["down"], idx
        out[f"L{l}_mlp_up_bias"] = sd[pre + "fc1.bias"][idx]
        out[f"L{l}_mlp_down_bias"] = sd[pre + "fc2.bias"]
        # type II
        r_qk = O.head_rank(hd, keep[l], rope=False)
        qk, mask = O.qk_layer(sd[pre + "self_attn.q_proj.weight"], sd[pre + "self_attn.k_proj.weight"], c_q,
                              c_k, H, H, hd, r_qk, "opt", 0.0, b_q=sd[pre + "self_attn.q_proj.bias"],
                              b_k=sd[pre + "self_attn.k_proj.bias"])
        out[f"L{l}_qk_q_proj"], out[f"L{l}_qk_k_proj"] = qk["q_proj"], qk["k_proj"]
        out[f"L{l}_qk_q_bias"], out[f"L{l}_qk_k_bias"] = qk["q_bias"], qk["k_bias"]
        out[f"L{l}_qk_mask"] = mask
        # type III
        r_vo = min(O.head_rank(hd, keep[l], rope=False, clamp_to_head=False), hd)
        w_o = sd[pre + "self_attn.out_proj.weight"]
        vo, v64, o64 = O.vo_layer(sd[pre + "self_attn.v_proj.weight"], w_o, c_x, H, H, hd, r_vo, ridge_vo)
        out[f"L{l}_vo_v64"], out[f"L{l}_vo_o64"] = v64, o64
        fold = w_o.astype(np.float32) @ sd[pre + "self_attn.v_proj.bias"].astype(np.float32) \
            + sd[pre + "self_attn.out_proj.bias"].astype(np.float32)
        out[f"L{l}_vo_o_bias"] = O.to_bf16(fold)
        out[f"L{l}_ranks"] = np.array([rank, r_qk, r_vo])
    return out
