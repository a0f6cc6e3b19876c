#!/usr/bin/env python
"""After `gpurun -- 'bash tools/r2_final.sh'`: turn gpurun_out/final_* into the files kept under profiles/
(launch lists, ncu summary of the last capture of every kernel, traffic.json, bench lines).  Runs here, no GPU."""
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def run(*a):
    return subprocess.run(a, capture_output=True, text=True, cwd=ROOT).stdout


def main():
    for src, cmd, dst in (("final_launches_bench.csv", "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-verify",
                           "launch_list_bench.txt"),
                          ("final_launches_shard8.csv", "python bench.py --steps 2 --warmup 3 --emulate-shard 8 --no-e2e --no-cpu-baseline",
                           "launch_list_shard8.txt")):
        open(os.path.join(PROF, "%s_%s" % (TAG, dst)), "w").write(
            run(sys.executable, "tools/launch_list.py", os.path.join(OUT, src), cmd))
    txt = run(sys.executable, "tools/ncu_summary.py", os.path.join(OUT, "final_prof_step.ncu-rep"))
    secs = [s for s in txt.split("=" * 100) if s.strip()]
    raw, src, order = {}, {}, []
    for s in secs:
        name = s.strip().split("\n")[0]
        key = re.sub(r"\((?:int|bool)\)", "", name).split("(")[0].strip()
        if "gpu__time_duration" in s:
            if key not in raw:
                order.append(key)
            raw[key] = s
        else:
            src[key] = s
    out = ["# ncu --set full --clock-control none --import-source on, tools/profile_r2.py (PASSES=1 SCATTER=1 NT=100): one launch of",
           "# every kernel of the path at the benchmark size (1 240 000 rays, 256x256x128); tools/ncu_summary.py of",
           "# gpurun_out/final_prof_step.ncu-rep (tools/r2_final.sh, tools/collect_profiles.py).  The step's kernels: quads_list,",
           "# prepared_forward, residual, prepared_adjoint, finish_gradient; the others are the alternatives (stateless sweep /",
           "# run-aggregated adjoint, binned operator).", ""]
    for k in order:
        out += ["=" * 100, raw[k].strip("\n")]
        if k in src:
            out += ["  -- source page --", "\n".join(src[k].strip("\n").split("\n")[1:])]
    open(os.path.join(PROF, "%s_ncu_step_summary.txt" % TAG), "w").write("\n".join(out) + "\n")

    def grab(key, metric):
        m = re.search(re.escape(metric) + r"\s+([0-9.]+)\s+(\w+)", raw[key])
        return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[m.group(2)]

    def tot(key):
        return grab(key, "dram__bytes_read.sum") + grab(key, "dram__bytes_write.sum")

    def find(prefix):
        return [k for k in raw if k.replace("void ", "").startswith(prefix)][0]

    tj = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch from ONE ncu --set full capture of the path's "
                      "kernels at the benchmark size (NT=100: 1 240 000 rays, 256x256x128 grid): profiles/%s_ncu_step_summary.txt "
                      "(gpurun_out/final_prof_step.ncu-rep, tools/r2_final.sh). bench.py only quotes an entry whose `rays` "
                      "equals the rays of the run." % TAG}
    for name, prefix, extra in (("prepared_forward", "prepared_forward_kernel", ()),
                                ("prepared_adjoint", "prepared_adjoint_kernel", ("finish_gradient_kernel",)),
                                ("binned_adjoint", "backproject_wruns_kernel", ("backproject_combine_short_kernel", "backproject_combine_kernel")),
                                ("ray_sweep_forward", "ray_sweep_kernel<0", ()),
                                ("ray_sweep_adjoint_scatter", "adjoint_runs_kernel", ())):
        k = find(prefix)
        b = tot(k) + sum(tot(find(e)) for e in extra)
        tj[name] = {"kernel": k.replace("void ", "") + (" (+ %s)" % ", ".join(extra) if extra else ""), "rays": 1240000,
                    "dram_bytes": int(b), "source": "profiles/%s_ncu_step_summary.txt" % TAG}
    json.dump(tj, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    for s, d in (("final_bench_n1.json", "bench_n1.json"), ("final_bench_ref.json", "bench_n1_reference_arm.json"),
                 ("final_bench_emul8.json", "bench_shard8_emulated.json"), ("final_kernel_bench.json", "kernel_bench.json"),
                 ("final_inv_c4.json", "inversion_lbfgs_c4.json"), ("final_inv_c2.json", "inversion_lbfgs_c2.json")):
        if os.path.exists(os.path.join(OUT, s)):
            shutil.copy(os.path.join(OUT, s), os.path.join(PROF, "%s_%s" % (TAG, d)))
    d = json.loads([l for l in open(os.path.join(PROF, "%s_bench_n1.json" % TAG)) if l.startswith("{")][-1])
    print("bench n1:", d["ms_per_step"], d["pass_frac_of_hbm_roofline"], d["roofline"], d["e2e"]["ms_per_step"],
          d["e2e_active_voxels"]["ms_per_step"], d["cpu_baseline"]["value"])
    print({k: round(v["ms"], 4) for k, v in d["kernels"].items()})
    for k, v in tj.items():
        if isinstance(v, dict):
            print(k, v["dram_bytes"] / 1e9)


if __name__ == "__main__":
    main()
