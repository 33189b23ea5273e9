"""Timeline of one step of the sharded loss head: which kernels and collectives run, on which stream, and how much of the
step the compute stream sits idle (= exposed communication + host glue).  No nsys in this image: the CUPTI trace of
torch.profiler is parsed instead.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port 29577 \
        tools/timeline.py [--global-batch 32768] > gpurun_out/timeline_Wgpu.txt
"""
import argparse
import json
import os
import sys
import tempfile

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--global-batch", type=int, default=32768)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--free-run", action="store_true",
                    help="no barrier / synchronize between the profiled steps (what the bench times): the step analysed is the "
                         "middle one, from the end of its predecessor's last kernel, so host run-ahead and inter-step gaps count")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    inst = bench.Instrument()
    flush = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
    head = bench.LossHead(args.global_batch, dev, rank, world, inst, flush)
    for _ in range(3):
        head.step(False)
    head.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        if args.free_run:
            args.steps = max(args.steps, 3)
            head.barrier()
            for _ in range(args.steps):
                head.step(False)
            torch.cuda.synchronize()
        else:
            for _ in range(args.steps):
                head.barrier()
                head.step(False)
                torch.cuda.synchronize()
    if rank == 0:
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, "trace.json")
            prof.export_chrome_trace(path)
            trace = json.load(open(path))
        ev = [e for e in trace["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
        ev.sort(key=lambda e: e["ts"])
        by_stream = {}
        for e in ev:
            by_stream.setdefault(e["args"].get("stream"), []).append(e)
        main_stream = max(by_stream, key=lambda s: sum(e["dur"] for e in by_stream[s]))
        # split into steps at the largest gaps of the main stream (the barrier + synchronize between steps)
        ms = by_stream[main_stream]
        gaps = sorted(range(1, len(ms)), key=lambda k: ms[k]["ts"] - (ms[k - 1]["ts"] + ms[k - 1]["dur"]), reverse=True)[:args.steps - 1]
        bounds = [0] + sorted(gaps) + [len(ms)]
        step_ev = ms[bounds[-2]:bounds[-1]]                    # the last profiled step
        t0 = step_ev[0]["ts"]
        if args.free_run:
            # every step launches the same kernels: split by count, take the middle step, start the clock where the step before ended
            # (a few kernels before the first and after the last step belong to the barrier / the profiler: find the trim
            #  that leaves `steps` identical name sequences)
            names = [e["name"] for e in ms]
            found = None
            for lead in range(0, 12):
                for trail in range(0, 12):
                    n = len(ms) - lead - trail
                    if n <= 0 or n % args.steps:
                        continue
                    per = n // args.steps
                    if all(names[lead + k * per:lead + (k + 1) * per] == names[lead:lead + per] for k in range(1, args.steps)):
                        found = (lead, per)
                        break
                if found:
                    break
            assert found, "no trim of the %d compute-stream kernels gives %d identical steps" % (len(ms), args.steps)
            lead, per = found
            mid = args.steps // 2
            step_ev = ms[lead + mid * per:lead + (mid + 1) * per]
            t0 = ms[lead + mid * per - 1]["ts"] + ms[lead + mid * per - 1]["dur"]
        t1 = step_ev[-1]["ts"] + step_ev[-1]["dur"]
        busy = sum(e["dur"] for e in step_ev)
        print("world size %d, global batch %d, rank 0, %s of %d profiled steps%s" % (
            world, args.global_batch, "middle" if args.free_run else "last", args.steps,
            " (free-running: no barrier or synchronize between steps)" if args.free_run else ""))
        print("step window on the compute stream (stream %s): %.3f ms, kernels busy %.3f ms, idle %.3f ms" % (
            main_stream, (t1 - t0) / 1e3, busy / 1e3, (t1 - t0 - busy) / 1e3))
        print("\nidle gaps of the compute stream > 20 us (what ran before / after, and what the other streams did meanwhile):")
        others = [e for s, lst in by_stream.items() if s != main_stream for e in lst if t0 <= e["ts"] <= t1]
        for a, b in zip(step_ev, step_ev[1:]):
            g0, g1 = a["ts"] + a["dur"], b["ts"]
            if g1 - g0 > 20:
                inside = [o for o in others if o["ts"] < g1 and o["ts"] + o["dur"] > g0]
                names = ", ".join("%s (%.0f us, stream %s)" % (o["name"][:48], o["dur"], o["args"].get("stream")) for o in inside) or "-"
                print("  %7.1f us  after %-40s before %-40s | other streams: %s" % (g1 - g0, a["name"][:40], b["name"][:40], names))
        print("\nkernels of the step by total time (compute stream):")
        agg = {}
        for e in step_ev:
            k = e["name"][:70]
            agg.setdefault(k, [0, 0.0])
            agg[k][0] += 1
            agg[k][1] += e["dur"]
        for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            print("  %9.3f ms  x%-3d %s" % (d / 1e3, n, k))
        print("\nother streams inside the step window (collectives, copies):")
        agg = {}
        for o in others:
            k = (o["name"][:70], o["args"].get("stream"))
            agg.setdefault(k, [0, 0.0, 0.0])
            agg[k][0] += 1
            agg[k][1] += o["dur"]
            # overlapped with compute-stream kernels?
            ov = sum(max(0, min(o["ts"] + o["dur"], e["ts"] + e["dur"]) - max(o["ts"], e["ts"])) for e in step_ev)
            agg[k][2] += ov
        for (k, s), (n, d, ov) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("  %9.3f ms  x%-3d stream %-4s %5.1f %% hidden behind compute kernels   %s" % (d / 1e3, n, s, 100.0 * ov / max(d, 1e-9), k))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
