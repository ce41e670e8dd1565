#!/usr/bin/env python3
"""Turns the ncu outputs gathered on the GPU box (gpurun_out/) into the committed summaries under
profiles/. Usage: python tools/summarise_profiles.py r01"""
import csv, os, subprocess, sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def launch_table(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    by = OrderedDict()
    for r in rows:
        by.setdefault((r[0], r[4].split("(")[0].strip()), {})[r[-3]] = float(r[-1].replace(",", ""))
    ids = list(by.keys())
    # one steady-state frame: from one single-pass dice launch to the next (the first frame of a renderer
    # sizes its buffers with the ordered two-pass dice and is skipped)
    starts = [i for i, k in enumerate(ids) if "k_dice_stream" in k[1]] or [i for i, k in enumerate(ids) if "k_dice<0>" in k[1]]
    s, e = (starts[-2], starts[-1]) if len(starts) > 1 else (starts[0], len(ids))
    return [(k[1], by[k]) for k in ids[s:e]]


lines = [f"# {tag} — ncu launch lists, one frame per scene", "",
         "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum "
         "--clock-control none python tools/stage_times.py <scene>` on a B200. Per-launch times are cold-cache and "
         "serialised: compare shares, not absolutes (the CUDA-event stage times of the same build are in the bench line).", ""]
for scene in ("random100k", "tiger4k"):
    path = os.path.join(ROOT, "gpurun_out", f"{tag}_launches_{scene}.csv")
    if not os.path.exists(path):
        continue
    table = launch_table(path)
    tot = sum(m.get("gpu__time_duration.sum", 0) for _, m in table) / 1000
    lines += [f"## {scene}", "", "| kernel | us | share | DRAM read MB | DRAM write MB | warp instr (M) |", "|---|---:|---:|---:|---:|---:|"]
    for name, m in table:
        t = m.get("gpu__time_duration.sum", 0) / 1000
        lines.append(f"| `{name}` | {t:.1f} | {100 * t / tot:.1f}% | {m.get('dram__bytes_read.sum', 0) / 1e6:.1f} | "
                     f"{m.get('dram__bytes_write.sum', 0) / 1e6:.1f} | {m.get('smsp__inst_executed.sum', 0) / 1e6:.1f} |")
    lines += [f"| **total** | {tot:.1f} | | | | |", ""]
open(os.path.join(out_dir, f"{tag}_launches.md"), "w").write("\n".join(lines))

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]
import hashlib, json


def raw_metrics(rep, launch=0):
    """{metric: (unit, value)} of one launch of an .ncu-rep (page raw)."""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3 + launch:
        return {}
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2 + launch])}


def number(text):
    try:
        return float(text.replace(",", ""))
    except ValueError:
        return 0.0


SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
lines = [f"# {tag} — `ncu --set full` of the fill + tile stage: k_tile_solid + k_tile_alpha", "",
         "One steady-state frame per scene; the two kernels run back to back on the renderer's stream (profiles are "
         "serialised and cold-cache: compare shares and counters, the CUDA-event stage time is in the bench line).", ""]
traffic = {}
for scene in ("random100k", "tiger4k"):
    for kernel in ("k_tile_solid", "k_tile_alpha"):
        rep = os.path.join(ROOT, "gpurun_out", f"{tag}_{kernel}_{scene}.ncu-rep")
        if not os.path.exists(rep):
            continue
        m = raw_metrics(rep)
        lines += [f"## {scene} — {kernel}", "", "| metric | unit | value |", "|---|---|---:|"]
        for k in KEYS + ["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                         "l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed",
                         "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_tex_mem_texture.sum"]:
            if k in m:
                lines.append(f"| `{k}` | {m[k][0]} | {m[k][1]} |")
        lines.append("")
        if scene == "random100k":
            traffic[kernel] = sum(number(m[k][1]) * SCALE.get(m[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
open(os.path.join(out_dir, f"{tag}_fill_tile_ncu.md"), "w").write("\n".join(lines))
if len(traffic) == 2:
    # bench.py reports this as roofline.traffic — only while composite.cu is the file that was profiled
    src = open(os.path.join(ROOT, "pathfinder_b200", "csrc", "composite.cu"), "rb").read()
    json.dump({"scene": "random100k@8192", "dram_bytes_per_frame": int(sum(traffic.values())),
               "per_kernel": {k: int(v) for k, v in traffic.items()}, "composite_cu_sha256": hashlib.sha256(src).hexdigest()},
              open(os.path.join(out_dir, f"{tag}_fill_tile_traffic.json"), "w"), indent=1)

# bin (count pass): the L2 atomic / reduction evidence the north star names
reps = [(name, os.path.join(ROOT, "gpurun_out", f"{tag}_k_bin_{name}_random100k.ncu-rep")) for name in ("count", "emit")]
if all(os.path.exists(rep) for _, rep in reps):
    BIN_KEYS = KEYS + ["lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_requests_op_atom.sum", "lts__t_requests_op_red.sum",
                       "l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum",
                       "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
    lines = [f"# {tag} — `ncu --set full` of bin, one steady-state frame of random100k@8192: k_bin<1> (count pass) and k_bin<0> (emit pass after the z-cull)", "",
             "Every fill and every backdrop change of the count pass is one 32-bit L2 reduction (`RED`, no return value) on the "
             "tile word; the emit pass claims its slot with one `ATOMG` per surviving fill.", ""]
    for name, rep in reps:
        m = raw_metrics(rep, 0)
        if not m:
            continue
        lines += [f"## {name} pass: `{m.get('Kernel Name', ('', '?'))[1]}`", "", "| metric | unit | value |", "|---|---|---:|"]
        for k in BIN_KEYS:
            if k in m:
                lines.append(f"| `{k}` | {m[k][0]} | {m[k][1]} |")
        lines.append("")
    open(os.path.join(out_dir, f"{tag}_bin_ncu.md"), "w").write("\n".join(lines))

# SASS evidence: memory / texture / atomic / shuffle mnemonics per kernel.
so = os.path.join(ROOT, "pathfinder_b200", "libpf_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kernels, cur = OrderedDict(), None
for ln in sass.splitlines():
    if "Function :" in ln:
        cur = ln.split("Function :")[1].strip()
        kernels[cur] = {}
    elif cur and "/*" in ln and ";" in ln:
        body = ln.split("*/")[1].strip() if "*/" in ln else ""
        op = body.split()[0] if body else ""
        if op.startswith("@"):
            op = body.split()[1] if len(body.split()) > 1 else ""
        for pat in ("LDG.E.128", "LDG.E.64", "LDG", "STG.E.128", "STG.E.64", "STG", "TEX", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "MATCH",
                    "LDS", "STS", "LDL", "STL", "BAR", "FFMA2", "FMUL2", "FFMA", "MUFU", "HMMA", "UTC", "UTMA"):
            if op.startswith(pat):
                kernels[cur][pat] = kernels[cur].get(pat, 0) + 1
                break
lines = [f"# {tag} — SASS mnemonic counts per kernel (`cuobjdump -sass libpf_cuda.so`, sm_100a)", "",
         "No tensor-core (`HMMA`/`UTC*MMA`) or TMA (`UTMA*`) instructions are expected: nothing on this path is a dense "
         "contraction or a bulk tile copy (BASELINE.json north_star). `FFMA2`/`FMUL2` are sm_100's packed f32x2 operations (the compositing blend). `LDL`/`STL` = local memory (the dice deep-recursion fallback).", "",
         "| kernel | " + " | ".join(["LDG.E.128", "LDG.E.64", "LDG", "STG.E.128", "STG", "TEX", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "LDS", "STS", "LDL", "STL", "BAR", "FFMA2", "FMUL2", "FFMA", "HMMA", "UTC", "UTMA"]) + " |",
         "|---|" + "---:|" * 22]
import re
for name, c in kernels.items():
    short = re.sub(r"^_ZN2pf\d+", "", name)[:40]
    lines.append(f"| `{short}` | " + " | ".join(str(c.get(k, 0)) for k in ["LDG.E.128", "LDG.E.64", "LDG", "STG.E.128", "STG", "TEX", "ATOMG", "RED", "ATOMS", "SHFL", "VOTE", "LDS", "STS", "LDL", "STL", "BAR", "FFMA2", "FMUL2", "FFMA", "HMMA", "UTC", "UTMA"]) + " |")
open(os.path.join(out_dir, f"{tag}_sass.md"), "w").write("\n".join(lines) + "\n")

# SASS excerpts: the inner loops themselves, not only counts.
def function_sass(match):
    out, on = [], False
    for ln in sass.splitlines():
        if "Function :" in ln:
            on = match in ln
        elif on and "/*" in ln and ";" in ln:
            out.append(ln.split("*/")[1].split(";")[0].strip() if ln.strip().startswith("/*") else ln.strip())
    return out


def excerpt(body, first_pat, last_pat, before=24, after=30):
    idx = [i for i, l in enumerate(body) if first_pat in l]
    if not idx:
        return []
    end = [i for i, l in enumerate(body) if last_pat in l and i >= idx[0]]
    lo, hi = max(0, idx[0] - before), min(len(body), (end[0] if end else idx[0]) + after)
    return body[lo:hi]


ex = [f"# {tag} — SASS excerpts (`cuobjdump -sass libpf_cuda.so`, sm_100a)", ""]
alpha = function_sass("k_tile_alphaILb0ELb0")
ex += ["## k_tile_alpha<false,false>: the per-fill loop (two fills per round: 2 x LDS.128 of the staged parameters each, "
       "window / t / y arithmetic, 2 x TEX.LL of the area LUT each, 8 FFMA + 8 integer adds into the coverage accumulators)", "", "```"]
ex += excerpt(alpha, "TEX.LL", "TEX.LL", before=40, after=40) + ["```", ""]
ex += ["## k_tile_alpha<false,false>: the blend of a mask over the 8 pixels of a lane (LDS.128 pixel, FFMA2 pairs, STS.128)", "", "```"]
ex += excerpt(alpha, "FFMA2", "FFMA2", before=6, after=40) + ["```", ""]
binc = function_sass("k_binILi1E")
ex += ["## k_bin<1> (count pass): add_fill's quantisation and the reduction on the tile word", "", "```"]
ex += excerpt(binc, "RED", "RED", before=40, after=12) + ["```", ""]
open(os.path.join(out_dir, f"{tag}_sass_excerpts.md"), "w").write("\n".join(ex) + "\n")
print("wrote", [f for f in os.listdir(out_dir) if f.startswith(tag)])
