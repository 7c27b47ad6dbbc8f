"""Summarise a MG_PROFILE_TIMELINE file (tools/gpu_timeline_type1.py): per call and lane the busy
time, and along the call the chain of launches that ended last ("critical path by hindsight":
starting from the launch that finishes last, step to the launch on any lane that ended closest
before this one started).
    python tools/timeline_summary.py gpurun_out/type1_timeline.csv [out.txt]"""
import collections
import sys


def load(path):
    reps, cur = [], None
    for line in open(path):
        line = line.strip()
        if line.startswith("#rep"):
            cur = collections.defaultdict(list)
            reps.append(cur)
            continue
        call, label, lane, a, b = line.split(",")
        cur[call].append((label, lane, float(a), float(b)))
    return reps


def summarise(call, spans, out):
    t_end = max(b for _, _, _, b in spans)
    lanes = collections.OrderedDict()
    for label, lane, a, b in spans:
        lanes.setdefault(lane, []).append((label, a, b))
    out.write(f"== {call}: {len(spans)} launches, {t_end:.3f} ms from the first launch to the last end\n")
    for i, (lane, ls) in enumerate(lanes.items()):
        busy = sum(b - a for _, a, b in ls)
        kinds = collections.Counter(l for l, _, _ in ls)
        top = ", ".join(f"{k} x{v}" for k, v in kinds.most_common(4))
        out.write(f"  lane {i}: {len(ls):4d} launches, busy {busy:7.3f} ms ({100 * busy / t_end:5.1f} %), "
                  f"first start {min(a for _, a, _ in ls):7.3f}, last end {max(b for _, _, b in ls):7.3f}: {top}\n")
    # hindsight critical path
    lane_idx = {lane: i for i, lane in enumerate(lanes)}
    cur = max(spans, key=lambda s: s[3])
    path = []
    while True:
        path.append(cur)
        prev = [s for s in spans if s[3] <= cur[2] + 1e-4 and s is not cur]
        if not prev:
            break
        cur = max(prev, key=lambda s: s[3])
    path.reverse()
    on_path = collections.defaultdict(float)
    gaps = 0.0
    for i, (label, lane, a, b) in enumerate(path):
        on_path[(lane_idx[lane], label)] += b - a
        if i:
            gaps += max(0.0, a - path[i - 1][3])
    out.write(f"  hindsight path: {len(path)} launches, {sum(on_path.values()):.3f} ms in kernels + {gaps:.3f} ms of gaps\n")
    for (lane, label), v in sorted(on_path.items(), key=lambda kv: -kv[1])[:12]:
        n = sum(1 for s in path if s[0] == label and lane_idx[s[1]] == lane)
        out.write(f"    lane {lane} {label:28s} {v:7.3f} ms in {n:4d} launches ({1e3 * v / n:6.1f} us each)\n")


if __name__ == "__main__":
    reps = load(sys.argv[1])
    out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
    for call, spans in reps[-1].items():
        summarise(call, spans, out)
