"""profiles/<name>.md from one GPU-box visit (tools/gpu_round.sh): the ncu launch list of the training step (one step = the launches between
two im2col kernels) and the `--set full` tables of a forward and a backward stretch.
    python tools/train_profile_md.py gpurun_out > profiles/r2_train_step_v2.md"""
import collections
import csv
import re
import sys


def short(n):
    n = re.sub(r"lpi::(\(anonymous namespace\)::)?", "", n)
    n = re.sub(r"<unnamed>::", "", n)
    n = re.sub(r"\(.*", "", n)
    return n.replace("void ", "")[:80]


def launch_table(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    seq = []
    for r in rows[1:]:
        try:
            seq.append((r[ki], float(r[vi].replace(",", "")) / 1000.0))
        except (ValueError, IndexError):
            pass
    idx = [i for i, s in enumerate(seq) if "im2col" in s[0]]
    a, b = idx[0], idx[1]
    agg = collections.OrderedDict()
    for n, t in seq[a:b]:
        c = agg.setdefault(short(n), [0, 0.0])
        c[0] += 1
        c[1] += t
    tot = sum(v[1] for v in agg.values())
    out = [f"One step = {b - a} launches, {tot:.0f} us serialised (cold caches: read the SHARES).\n",
           "| us | share | launches | avg us | kernel |", "|---|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if v[1] / tot < 0.001:
            continue
        out.append(f"| {v[1]:.0f} | {100 * v[1] / tot:.1f} % | {v[0]} | {v[1] / v[0]:.1f} | `{k}` |")
    return "\n".join(out)


def full_table(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    col = {k: hdr.index(k) for k in ("Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                                     "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                     "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size",
                                     "l1tex__m_xbar2l1tex_read_bytes.sum") if k in hdr}
    units = rows[1]
    out = ["| kernel | us | tensor pipe active % | DRAM read + write MB | DRAM % | L2 hit % | L2 -> SM MB | regs | grid |", "|---|---|---|---|---|---|---|---|---|"]

    def mb(r, k):
        v = float(r[col[k]].replace(",", ""))
        u = units[col[k]].lower()
        return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)

    for r in rows[2:]:
        g = lambda k: float(r[col[k]].replace(",", ""))
        us = g("gpu__time_duration.sum") * (1e-3 if units[col["gpu__time_duration.sum"]] in ("ns", "nsecond") else 1.0)
        x = f"{mb(r, 'l1tex__m_xbar2l1tex_read_bytes.sum'):.0f}" if "l1tex__m_xbar2l1tex_read_bytes.sum" in col else "-"
        out.append(f"| `{short(r[col['Kernel Name']])}` | {us:.1f} | {g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{mb(r, 'dram__bytes_read.sum'):.0f} + {mb(r, 'dram__bytes_write.sum'):.0f} | {g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.0f} | "
                   f"{g('lts__t_sector_hit_rate.pct'):.0f} | {x} | {int(g('launch__registers_per_thread'))} | {int(g('launch__grid_size'))} |")
    return "\n".join(out)


if __name__ == "__main__":
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    print("## Launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -s 1250 -c 850 python tools/bench_train.py --batch 64 --steps 2 --warmup 3`)\n")
    print(launch_table(f"{d}/train_launches.csv"))
    print("\n## ncu `--set full` over a forward stretch (`-k regex:gemm_pair|attn_|layernorm -s 470 -c 10`)\n")
    print(full_table(f"{d}/train_fwd_raw.csv"))
    print("\n## ncu `--set full` over a backward stretch (`-s 640 -c 12`)\n")
    print(full_table(f"{d}/train_bwd_raw.csv"))
