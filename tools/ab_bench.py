"""A/B two builds of libflowdec_b200.so on one GPU box (boxes differ by +-3 % under the power cap, so
variants are only comparable inside one gpurun call).

  here (CPU box):   python tools/ab_bench.py --build "-DFD_EPI_STATS_SMEM=0 -DFD_MBAR_HINT_NS=0"
                    -> flowdec_b200/lib_variant.so next to the default build (both travel with gpurun)
  on the GPU box:   python tools/ab_bench.py --run [--rounds 2] [bench.py flags...]
                    -> alternates default / variant `bench.py --steps 3 --warmup 3 --no-cpu-baseline` runs

Build-time switches that exist: FD_XF_LAYOUT (transform-warp placement), FD_EPI_STATS_SMEM (column statistics by
shared-memory transpose vs shuffle butterfly), FD_MBAR_HINT_NS (mbarrier suspend-time hint, 0 = polling loop),
FD_HALO_DEBUG (cycle counters + transform ablation modes for tools/halo_dbg.py).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANT = os.path.join(ROOT, "flowdec_b200", "lib_variant.so")


def main():
    args = sys.argv[1:]
    if args[:1] == ["--build"]:
        sys.path.insert(0, ROOT)
        from flowdec_b200.build import build
        build()
        print(build(defs=args[1].split(), lib_out=VARIANT))
        return
    if args[:1] != ["--run"]:
        sys.exit(__doc__)
    args = args[1:]
    rounds = 2
    if args[:1] == ["--rounds"]:
        rounds, args = int(args[1]), args[2:]
    for _ in range(rounds):
        for name, lib in (("default", ""), ("variant", VARIANT)):
            env = dict(os.environ, FD_LIB_PATH=lib)
            r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3",
                                "--no-cpu-baseline", "--no-extras"] + args, env=env, capture_output=True, text=True, timeout=900)
            try:
                d = json.loads(r.stdout.strip().splitlines()[-1])
                print(f"{name:8s} {d['value']:8.2f} audio-s/s  e2e {d['e2e']['value']:8.2f}  clk {d['clocks']['sm_mhz']:6.0f} MHz  "
                      f"conv frac {d['roofline']['frac']:.3f}")
            except Exception as e:  # noqa
                print(name, "FAILED", e, r.stderr[-400:])


if __name__ == "__main__":
    main()
