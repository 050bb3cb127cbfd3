#!/bin/bash
# Copies the round's ncu / trace evidence from gpurun_out/ (scratch) into profiles/ (tracked).  Usage: tools/collect_profiles.sh <tag>
set -e
cd "$(dirname "$0")/.."
tag=$1
cp gpurun_out/r1_bf16${tag}_launches.csv profiles/r1_bf16_launches.csv
python tools/launch_summary.py profiles/r1_bf16_launches.csv conv3x3_tc > profiles/r1_bf16_launches_summary.txt
ncu -i gpurun_out/r1_bf16${tag}_top.ncu-rep --page raw --csv 2>/dev/null > profiles/r1_conv_tc_top_ncu_raw.csv
ncu -i gpurun_out/r1_bf16${tag}_top.ncu-rep --page source --print-source cuda,sass --csv --launch-skip 0 --launch-count 1 2>/dev/null > /tmp/top_src.csv
python tools/ncu_lines.py /tmp/top_src.csv 40 > profiles/r1_conv_tc_top_lines.txt
python tools/ncu_roles.py /tmp/top_src.csv speech-denoising-diffusion-model-2_b200/csrc/conv_tc.cu >> profiles/r1_conv_tc_top_lines.txt
cp gpurun_out/r1_trace.txt profiles/r1_pipeline_trace.txt
cp gpurun_out/r1_umma_rate.txt profiles/r1_umma_rate.txt
[ -f gpurun_out/r1_stft_bench.txt ] && cp gpurun_out/r1_stft_bench.txt profiles/r1_stft_bench.txt
[ -f gpurun_out/r1_parity_report.txt ] && cp gpurun_out/r1_parity_report.txt profiles/r1_parity_report.txt
python - <<'PY'
import csv, json
rows = list(csv.reader(open("profiles/r1_conv_tc_top_ncu_raw.csv")))
hdr = rows[0]
r, w, t = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
units = rows[1]
def to_bytes(v, u):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
labels = ["conv:ups.14.block1", "conv:ups.14.block2"]      # -s 82 -c 2 of the second forward: conv launches 40, 41
out = {"batch": 64, "precision": "bf16act", "source": "ncu --set full, profiles/r1_conv_tc_top_ncu_raw.csv", "dram_bytes_per_launch": {}, "ncu_duration_us": {}}
for lab, row in zip(labels, rows[2:]):
    out["dram_bytes_per_launch"][lab] = to_bytes(row[r], units[r]) + to_bytes(row[w], units[w])
    out["ncu_duration_us"][lab] = float(row[t])
json.dump(out, open("profiles/top_kernel_traffic.json", "w"), indent=1)
print(out)
PY
