"""File what tools/evidence.sh left in gpurun_out/ under profiles/ (tracked): launch lists, bench lines, the
ncu --set full summaries (and the per-launch DRAM traffic bench.py quotes), probe / sweep tables, driver lines,
the host pipeline's timeline and SASS listings of the current build.  Runs here (no GPU): it only reads files
and calls `ncu -i` / `cuobjdump`."""
import csv
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

COPY = {
    "launches_bench.csv": f"{TAG}_launches_bench.csv", "launches_fp32.csv": f"{TAG}_launches_fp32.csv",
    "launches_bf16mix.csv": f"{TAG}_launches_bf16mix.csv", "bench_n1.json": f"{TAG}_bench_n1.json",
    "bench_reference.json": f"{TAG}_bench_reference.json", "encoder_bench.jsonl": f"{TAG}_encoder_bench.jsonl",
    "inference_bench.jsonl": f"{TAG}_inference_bench.jsonl", "sweep.jsonl": f"{TAG}_sweep_config5.jsonl",
    "host_pipeline_timeline.txt": f"{TAG}_host_pipeline_timeline.txt", "probe_n2.json": f"{TAG}_probe_encoder_n2.json",
    "probe_n16.json": f"{TAG}_probe_encoder_uniform.json", "bench_n2.json": f"{TAG}_bench_n2.json",
    "soc_step_n1.json": f"{TAG}_soc_n1.json", "level_breakdown.jsonl": f"{TAG}_level_breakdown.jsonl",
    "reference_stack_parity.txt": f"{TAG}_reference_stack_parity.txt", "reference_test_py.txt": f"{TAG}_reference_test_py.txt",
    "pytest_gpu.log": f"{TAG}_pytest_gpu.log",
}
for src, dst in COPY.items():
    if (OUT / src).exists():
        shutil.copy(OUT / src, PROF / dst)
        print("copied", dst)

# markdown table of the sweep
log = OUT / "sweep.log"
if log.exists():
    lines = log.read_text().splitlines()
    start = next((i for i, l in enumerate(lines) if l.startswith("| tokens")), None)
    if start is not None:
        (PROF / f"{TAG}_sweep_config5.md").write_text(
            "BASELINE.json configs[4] sweep + decoder shapes (tools/sweep.py --decoder; B200, CUDA-event medians, "
            "256 MB L2 flush between calls; N = 16 frames unless stated)\n\n" + "\n".join(lines[start:]) + "\n")
        print("wrote sweep table")

# ncu --set full summaries
traffic = {}
for rep, name in (("full_bf16mix", f"{TAG}_ncu_full_bf16mix_summary.txt"),
                  ("full_direct", f"{TAG}_ncu_full_direct_gather_summary.txt")):
    tmp = OUT / (rep + ".raw.csv")          # exported on the GPU box (the reports exceed gpurun's 64 MiB return)
    if not tmp.exists() and (OUT / (rep + ".ncu-rep")).exists():
        tmp.write_text(subprocess.run(["ncu", "-i", str(OUT / (rep + ".ncu-rep")), "--page", "raw", "--csv"],
                                      capture_output=True, text=True).stdout)
    if not tmp.exists():
        continue
    raw = tmp.read_text()
    txt = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), str(tmp)], capture_output=True, text=True).stdout
    (PROF / name).write_text(txt)
    print("wrote", name)
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[2:]:
        k = r[ki].split("<")[0].replace("void ", "").strip()
        traffic[k] = int(float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]])
if traffic:
    traffic["_source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, bf16 values + fp32 "
                          f"locations, N=16 (profiles/{TAG}_ncu_full_bf16mix_summary.txt)")
    (PROF / "dominant_kernel_traffic.json").write_text(json.dumps(traffic, indent=1))
    print("traffic", traffic)

# SASS of the kernels the bench launches
lib = ROOT / "neurips2023_soc_b200" / "libmsda_b200.so"
names = subprocess.run(["cuobjdump", "-sass", str(lib)], capture_output=True, text=True).stdout
want = {
    "fwd_tile_bf16": "msda_fwd_tile_kernelI13__nv_bfloat16fLi8ELi4ELi4ELb1ELb0ELi512E",
    "fwd_tile_f32": "msda_fwd_tile_kernelIffLi4ELi8ELi4ELb1ELb0ELi1024E",
    "bwd_sample_tile_bf16": "msda_bwd_sample_tile_kernelI13__nv_bfloat16fLi8ELi4ELi4ELb1ELb0ELi512E",
    "grad_value_walk_bf16": "msda_grad_value_walk_kernelI13__nv_bfloat16Li4ELi8E",
    "bin_rank_sort_f32": "msda_bin_rank_sort_kernelIfE",
    "grad_value_direct_f32": "msda_grad_value_direct_kernelIffLi4ELi8ELb1E",
    "bwd_bin_bf16": "msda_bwd_bin_kernelI13__nv_bfloat16Li8ELi4E",
    "add_layernorm_fwd_bf16": "msda_add_layernorm_fwd_kernelI13__nv_bfloat16E",
    "add_layernorm_bwd_bf16": "msda_add_layernorm_bwd_kernelI13__nv_bfloat16E",
}
blocks = names.split("\t\tFunction : ")
for tag, key in want.items():
    for b in blocks[1:]:
        if key in b.split("\n", 1)[0]:
            keep = [l for l in ("\t\tFunction : " + b).splitlines()]
            (PROF / f"{TAG}_sass_{tag}.txt").write_text("\n".join(keep) + "\n")
            print("sass", tag, len(keep), "lines")
            break
    else:
        print("sass: no match for", tag)
