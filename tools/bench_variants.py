"""Run bench.py in-process with backbone / conv variants (development tool)."""
import contextlib, io, json, os, runpy, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowdec_b200 import ops
import flowdec_b200.backbones.ncsnpp as N

halo, fuse = int(os.environ.get("FD_HALO", "1")), int(os.environ.get("FD_FUSE", "1"))
ops.HALO_TILES = bool(halo)
_orig = N.NCSNpp.__init__


def patched(self, *a, **k):
    _orig(self, *a, **k)
    self.fuse_gn_into_conv = bool(fuse)


N.NCSNpp.__init__ = patched
sys.argv = ["bench.py", "--steps", "3", "--warmup", "3", "--no-cpu-baseline", "--no-extras"] + sys.argv[1:]
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), run_name="__main__")
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print("halo", halo, "fuse", fuse, "ring", os.environ.get("FD_HALO_RING", "0"), "value", round(d["value"], 2), "ms",
      round(d["ms_per_step"], 1), "conv_ms", round(d["roofline"]["kernel_ms_per_step"], 1), "frac",
      round(d["roofline"]["frac"], 3), "whole", round(d["roofline"]["whole_step_frac"], 3), "clk", d["clocks"]["sm_mhz"])
