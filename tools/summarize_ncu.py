"""Turn ncu outputs from gpurun_out/ into the small, committed summaries under profiles/.

    python tools/summarize_ncu.py launches <launches.csv> <out.txt> [--title "..."]
    python tools/summarize_ncu.py full <report.ncu-rep> <out.json>
"""
import collections
import csv
import json
import re
import subprocess
import sys


def us(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(u, v)


OWN = ["bi_cosine_kernel", "finalize_sym_kernel", "scale_kernel", "split_planes_kernel", "potrf128_kernel",
       "trsm128_kernel", "select_k_kernel", "gather_rows_kernel", "gather_sym_kernel",
       "gather_cols_planes_kernel", "copy_ridge_kernel", "identity_kernel", "transpose_to_bf16_kernel",
       "transpose_bf16_kernel", "qk_select_kernel", "vo_factor_kernel", "vo_apply_v_kernel",
       "vo_apply_o_kernel", "diag_block_kernel", "ydiag_kernel", "rmsnorm_kernel", "swiglu_kernel",
       "rope_kernel", "rope_masked_kernel", "rmsnorm_masked_kernel", "ce_rows_kernel", "pack_upper_kernel",
       "gram64_kernel", "f32_to_f64_kernel", "copy_upper_ridge_kernel", "mark_kept_kernel",
       "sub_jitter_rows_kernel", "add_kept_transpose_kernel"]


def short_name(name: str) -> str:
    if "gemm_tn_pair_kernel" in name:
        return "mg::gemm_tn_pair_kernel"
    if "gemm_tn_kernel" in name:
        return "mg::gemm_tn_kernel" + ("<256>" if "<256>" in name else "<128>")
    for own in OWN:
        if re.search(r"\b" + own + r"\b", name):
            return "mg::" + own
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<.*", "", name)
    return name.split("::")[-1].strip()[:56] or "(lambda)"


def _bytes(row):
    v = float(row["Metric Value"].replace(",", ""))
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(row["Metric Unit"], 1.0)


def launches(path, out, title):
    """Per-kernel launch count, time and share; when the capture also holds dram__bytes_read/write,
    the DRAM bytes per launch and the DRAM GB/s they amount to (for the HBM-bound kernels)."""
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt, dram = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
    has_dram = False
    for r in rows:
        k = short_name(r["Kernel Name"])
        name = r.get("Metric Name", "gpu__time_duration.sum")
        if name == "gpu__time_duration.sum":
            tot[k] += us(r)
            cnt[k] += 1
        elif name.startswith("dram__bytes"):
            dram[k] += _bytes(r)
            has_dram = True
    T = sum(tot.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: {path}  (ncu --metrics gpu__time_duration.sum --clock-control none; "
                f"per-launch times are cold-cache and serialised: compare SHARES)\n")
        f.write(f"# {sum(cnt.values())} launches, {T / 1e3:.2f} ms of kernel time\n")
        extra = f" {'dram_MB/launch':>15s} {'dram_GB/s':>10s}" if has_dram else ""
        f.write(f"{'kernel':58s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}{extra}\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            line = f"{k:58s} {cnt[k]:8d} {v / 1e3:10.3f} {v / cnt[k]:10.1f} {100 * v / T:6.1f}%"
            if has_dram:
                line += f" {dram[k] / cnt[k] / 1e6:15.1f} {dram[k] / (v * 1e-6) / 1e9:10.0f}"
            f.write(line + "\n")


METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
]


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    res = []
    for r in body:
        d = {"kernel": short_name(r[hdr.index("Kernel Name")]), "grid": r[hdr.index("Grid Size")]}
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                d[m] = {"value": r[i], "unit": units[i]}
        res.append(d)
    json.dump({"source": rep, "kernels": res}, open(out, "w"), indent=1)


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        title = sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--title" else src
        launches(src, dst, title)
    else:
        full(src, dst)
