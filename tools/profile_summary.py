"""Builds profiles/README.md from the captured files (bench JSON, ncu launch list, ncu --set full raw page).

    python tools/profile_summary.py r02
reads  profiles/<tag>_bench.json, <tag>_launches.csv, <tag>_ncu_full_raw.csv (any may be missing)."""
import csv, io, json, pathlib, sys
from collections import defaultdict

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
P = pathlib.Path(__file__).resolve().parent.parent / "profiles"
out = [f"# Profiles `{tag}` (B200, sm_100a)", "",
       "The launch list and the live table are of `python bench.py --steps K --warmup W --no-cpu-baseline` (BASELINE configs[1]: 4096 x 1 s clips, "
       "mel + CNN); the full capture is of `python tools/stft_only.py full` (the same 4096 clips: image chain, MFCC chain, then the fused full "
       "ensemble), one launch per kernel.", ""]
INDEX = """## Files of this round

| file | what |
|---|---|
| `r02_bench.json`, `r02_bench_reference.json` | the driver's two arms for cfg 2 (`bench.py`, `bench.py --impl reference`) as run by the builder at HEAD |
| `r02_bench_cfg1.json`, `_cfg3.json`, `_cfg4.json`, `_cfg4_file.json`, `_cfg5_1gpu.json` | `bench.py --config 1 / 3 / 4 / 4 --cfg4-mode file / 5` on one B200 |
| `r02_cfg4_{1,2,4,8}gpu.json`, `r02_cfg4_file_8gpu.json` | BASELINE configs[3]: one hour of audio at 1 / 2 / 4 / 8 GPUs (phrases sharded; one contiguous file); the 8-GPU line re-run with the block-FFT YIN, 1 / 2 / 4 and the file mode mid-round |
| `r02_bench_2gpu.json` | cfg 2 on two GPUs at HEAD (`torchrun ... bench.py --gpus 2`) |
| `r02_cfg5_8gpu.json`, `r02_cfg3_8gpu.json`, `r02_bench_8gpu.json` | BASELINE configs[4] sweep with the CPU column, the full ensemble, and cfg 2, on 8 GPUs (mid-round kernels) |
| `r02_h2d_probe.json` | pure pinned H2D copies at 1 / 2 / 4 / 8 ranks (`tools/h2d_probe.py`): the end-to-end ceiling |
| `r02_parity.json` | error distribution of every stage over all 4096 bench clips (`tools/parity_report.py`) |
| `r02_launches.csv`, `r02_ncu_full_raw.csv` | ncu launch list and `--set full` raw page at HEAD |
| `r02_sass_tc.txt` | per-kernel counts of UTCHMMA / LDTM / UBLKCP / UTCBAR / packed-FP32 instructions in the shipped `libgat.so` |
| `traffic.json` | DRAM bytes per launch from the full capture (bench.py's `roofline.traffic`) |
| `r02_fma2_probe.txt`, `r02_topo_8gpu.txt` | FFMA vs FFMA2 throughput probe (`tests/gpu_probe/fma2_probe.cu`); `nvidia-smi topo -m` of the 8-GPU box |
| `r02_early/` | earlier states of this round (v1: round-1 kernels + new bench; v2: frame kernel before the packed twiddles) |
| `r01_*` | round 1 |
"""


def csv_rows(path):
    text = path.read_text(errors="replace")
    start = text.index('"ID"')
    return list(csv.DictReader(io.StringIO(text[start:])))


lp = P / f"{tag}_launches.csv"
if lp.exists():
    rows = csv_rows(lp)
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        name = r["Kernel Name"].split("(")[0]
        agg[name][0] += 1; agg[name][1] += ms
    tot = sum(v[1] for v in agg.values()) or 1.0
    out += ["## ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised)", "",
            "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| {k[:90]} | {n} | {ms:.3f} | {ms / tot:.3f} |")
    out.append("")

bp = P / f"{tag}_bench.json"
if bp.exists():
    d = json.loads(bp.read_text())
    out += [f"## Live CUDA-event profile inside bench.py (same command, `kernels` array of profiles/{tag}_bench.json)", "",
            "| kernel | launches/step | ms/step | share | achieved | of measured peak |", "|---|---|---|---|---|---|"]
    for k in d["kernels"]:
        out.append(f"| {k['kernel']} | {k['launches_per_step']:g} | {k['ms_per_step']:.3f} | {k['share']:.3f} | {k['achieved']:.1f} {k['unit']} | {k['frac']:.3f} |")
    e = d["e2e"]
    out += ["", f"Step {d['ms_per_step']:.2f} ms -> {d['value']:.0f} audio-s/s resident, {e['value']:.0f} audio-s/s end to end "
            f"(host float32 buffers, {e['ms_per_step']:.2f} ms per step)" +
            (f", {d['e2e_pcm16']['value']:.0f} with PCM_16 host buffers" if "e2e_pcm16" in d else "") + f"; clocks {d['clocks']}."]
    if "fp32" in d.get("roofline", {}):
        f = d["roofline"]["fp32"]
        out.append(f"STFT kernel against the FP32-FMA roof measured on the same device: {f['achieved']:.1f} of {f['peak']:.1f} TFLOP/s = {f['frac']:.2f}.")
    out.append("")

fp = P / f"{tag}_ncu_full_raw.csv"
if fp.exists():
    rows = csv_rows(fp)
    cols = [("grid", "launch__grid_size"), ("duration", "gpu__time_duration.sum"), ("dram read", "dram__bytes_read.sum"),
            ("dram write", "dram__bytes_write.sum"), ("tensor pipe active %", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            ("issue active %", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
            ("fma pipe %", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            ("lsu pipe %", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("regs", "launch__registers_per_thread"),
            ("smem wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("smem conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
            ("L1 data pipe %", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed"), ("warp instructions", "smsp__inst_executed.sum")]
    units = rows[0] if rows and not rows[0].get("ID", "").isdigit() else {}
    out += [f"## ncu --set full (one launch each; `profiles/{tag}_ncu_full_raw.csv` holds every metric)", "",
            "| kernel | " + " | ".join(c for c, _ in cols) + " |", "|---|" + "---|" * len(cols)]
    for r in rows:
        if not r.get("ID", "").isdigit():
            continue
        vals = []
        for _, key in cols:
            hit = [k for k in r if k.endswith(key)]
            vals.append((r[hit[0]] + " " + units.get(hit[0], "")).strip() if hit else "")
        out.append(f"| {r['Kernel Name'].split('(')[0][:60]} | " + " | ".join(vals) + " |")
    out.append("")
yp = P / f"{tag}_yin_ncu_full_raw.csv"
if yp.exists():
    r = [x for x in csv_rows(yp) if x.get("ID", "").isdigit()][0]
    g = lambda key: r[[k for k in r if k.endswith(key)][0]]
    out += [f"## YIN kernel (`tools/yin_only.py`, 4096 x 1 s clips; `profiles/{tag}_yin_ncu_full_raw.csv`)", "",
            f"`yin_kernel<15>`: {float(g('gpu__time_duration.sum')):.2f} ms per launch (3.73 ms before the batched refill and the bounds-check-free "
            f"inner loop); issue slots {float(g('sm__issue_active.avg.pct_of_peak_sustained_elapsed')):.0f} %, FMA pipe "
            f"{float(g('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active')):.0f} %, FP64 pipe "
            f"{float(g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active')):.1f} %, {g('launch__registers_per_thread')} registers.", ""]
out.append(INDEX)
(P / "README.md").write_text("\n".join(out) + "\n")
print("\n".join(out)[:3000])
