import os, subprocess, sys, tempfile
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from refrun import HARNESS, ROOT, spec_for
from tribs_b200.synth import write_basin
d = tempfile.mkdtemp(prefix="tribshybp_")
write_basin(spec_for("shallow", 40, runtime_h=8.0, seed=4, tributary_every=8), d)
lib = os.path.join(ROOT, "tribs_b200", "libtribs_gpu.so")
out = os.path.join(d, "gpu"); os.makedirs(out)
env = dict(os.environ, TRIBS_DEBUG="1")
r = subprocess.run([HARNESS, "basin.in", "-K", "--no-basin", "--max-steps", "32", "--out", out, "--gpu", lib, "--gpu-ranks", "2", "--gpu-batch", "16"],
                   cwd=d, capture_output=True, text=True, timeout=600, env=env)
print("rc", r.returncode)
print(r.stdout[-1500:])
print("STDERR", r.stderr[-4000:])
