"""Turn the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.
    python profiles/summarize.py r01
Needs `ncu` (reads .ncu-rep files without a GPU)."""
import collections
import csv
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size", "sm__cycles_elapsed.avg.per_second",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "launch__block_size",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__m_l1tex2xbar_write_bytes_mem_dshared.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "launch__waves_per_multiprocessor",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max", "launch__occupancy_cluster_max_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    agg = collections.OrderedDict()
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        v = v / 1e3 if r["Metric Unit"] == "ns" else (v * 1e3 if r["Metric Unit"] == "ms" else v)
        a = agg.setdefault(r["Kernel Name"].split("(")[0][:70], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as fh:
        fh.write("# ncu --metrics gpu__time_duration.sum --clock-control none : %d launches, %.1f us total (cold-cache, serialised)\n" % (len(rows), tot))
        fh.write("# %10s %6s %5s  kernel\n" % ("us", "share", "n"))
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            fh.write("%12.1f %5.1f%% %5d  %s\n" % (t, 100 * t / tot, c, k))


def raw(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    hdr, units = r[0], r[1]
    if len(r) < 3:
        raise RuntimeError("no launches in " + rep)
    with open(out, "w") as fh:
        for vals in r[2:]:          # one block per captured launch
            d = dict(zip(hdr, vals))
            fh.write("# %s : kernel %s  grid %s  cluster %s\n" % (rep, d.get("Kernel Name", "?"), d.get("launch__grid_size", "?"),
                                                                 d.get("launch__cluster_size", "?")))
            for h, u, v in zip(hdr, units, vals):
                if h in KEYS:
                    fh.write("%-95s %18s %s\n" % (h, v, u))
            fh.write("\n")


launches("gpurun_out/launches_%s.csv" % tag, "profiles/launches_%s.txt" % tag)
import os
for k in ("bwd", "bwd_e", "fwd", "colgrad", "ema", "retrieval", "pooler"):
    if not os.path.exists("gpurun_out/prof_%s_%s.ncu-rep" % (k, tag)):
        continue
    try:
        raw("gpurun_out/prof_%s_%s.ncu-rep" % (k, tag), "profiles/ncu_%s_%s.txt" % (k, tag))
    except Exception as e:  # noqa
        print("skip", k, e)
