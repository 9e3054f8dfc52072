"""Turns the raw ncu exports under gpurun_out/ into the small tracked summaries under profiles/ (round 1, final build; the v16 files of the same names were produced by an earlier revision of this script)."""
import collections
import csv
import json
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, SRC = ROOT / "profiles", ROOT / "gpurun_out"


def launch_table(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("(anonymous namespace)::", "")
        us = float(r["Metric Value"].replace(",", "")) / 1000.0
        tot += us
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    with open(dst, "w") as f:
        f.write("kernel,launches,total_us,share\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write('"%s",%d,%.1f,%.4f\n' % (k, v[0], v[1], v[1] / tot))
        f.write('"TOTAL",%d,%.1f,1.0\n' % (len(rows), tot))
    return tot


def full_summary(src, dst, traffic_json, batch, bench_batch, alg_bytes_b16):
    rows = list(csv.reader(open(src)))
    hdr, data = rows[0], rows[2:]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "launch__grid_size", "smsp__inst_executed.sum"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx] + ["units: us / MB / MB / % ..."])
        for d in data:
            w.writerow([d[i][:60] for i in idx])
    name_i, dur_i = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    rd_i, wr_i = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    tp_i = hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed")
    k3 = [d for d in data if "conv_tc_kernel<3" in d[name_i] or "conv_row_kernel" in d[name_i] or "conv_rowg_kernel<3" in d[name_i]]
    dur = sum(float(d[dur_i]) for d in k3)
    dram = sum(float(d[rd_i]) + float(d[wr_i]) for d in k3) * 1e6
    tens = sum(float(d[tp_i]) * float(d[dur_i]) for d in k3) / dur
    json.dump({
        "source": "ncu --set full --clock-control none, conv_tc_kernel<3,*> + conv_row_kernel + conv_rowg_kernel<3,*>, the %d launches of one DEP-UResNet "
                  "forward at batch %d (profiles/%s)" % (len(k3), batch, dst.name),
        "launches": len(k3), "dram_bytes_total_batch%d" % batch: dram,
        "algorithmic_bytes_total_batch%d" % batch: alg_bytes_b16,
        "traffic_over_algorithmic": dram / alg_bytes_b16,
        "dram_bytes_per_launch": dram / len(k3) * bench_batch / batch,
        "note": "dram_bytes_per_launch = batch-%d capture scaled x%d to the bench batch of %d (traffic is proportional "
                "to the slice count; at batch %d part of the inter-layer traffic is absorbed by the 126 MB L2)"
                % (batch, bench_batch // batch, bench_batch, batch),
        "time_weighted_tensor_pipe_pct_batch%d_cold" % batch: tens,
    }, open(traffic_json, "w"), indent=1)


def round2(tag="r02", commit=None):
    """Round 2: launch list of one batch-64 forward and --set full on every convolution launch of it (scripts/r2_run4.sh)."""
    t = launch_table(SRC / "r02_infer_launches_b64_raw.csv", OUT / ("%s_infer_launches_b64.csv" % tag))
    print("inference forward b64: %.1f us over all launches (cold, serialised)" % t)
    full_summary(SRC / "r02_conv_full_raw_b64.csv", OUT / ("%s_conv_ncu_full_summary_b64.csv" % tag),
                 OUT / "ncu_traffic_r02.json", 64, 64, 64 * 100597760.0)
    d = json.load(open(OUT / "ncu_traffic_r02.json"))
    d["commit"] = commit
    d["algorithmic_bytes_definition"] = (
        "every input / output / FiLM-residual tensor of the 20 3x3 launches moved once, bf16 (fp32 for the 4-channel head "
        "output): 100 597 760 B per slice = the figure depgan_profile_end reports to bench.py (6.44 GB per 64-slice step); "
        "round 1's 8.92 GB counted the decoder inputs once per source AND once concatenated")
    json.dump(d, open(OUT / "ncu_traffic_r02.json", "w"), indent=1)


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "r02":
        round2("r02" if len(sys.argv) < 4 else sys.argv[3], sys.argv[2] if len(sys.argv) > 2 else None)
        sys.exit(0)
    # final build of round 1 (programmatic dependent launch): launch list and --set full, both at the bench batch of 64
    t = launch_table(SRC / "infer_launches_b64_final.csv", OUT / "r01_final_infer_launches_b64.csv")
    print("inference forward b64: %.1f us over all launches (cold, serialised)" % t)
    if (SRC / "train_launches_v25.csv").exists():
        t = launch_table(SRC / "train_launches_v25.csv", OUT / "r01_final_train_launches_b32.csv")
        print("train iteration b32: %.1f us" % t)
    # algorithmic bytes of the 20 3x3 launches: 2 230 382 600 at batch 16 x 4
    full_summary(SRC / "infer_conv_tc_full_raw_b64.csv", OUT / "r01_final_conv_tc_ncu_full_summary_b64.csv",
                 OUT / "ncu_traffic_r01.json", 64, 64, 4 * 2230382600.0)
