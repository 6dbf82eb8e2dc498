"""Runs the small convolution / encoder / explainer GPU tests against the bounds-checking debug build
(liblrpcap_dbg.so: -DLRPCAP_DEBUG_BOUNDS, every epilogue address checked against the logical size of its tensor; a
violation prints the site and traps).  compute-sanitizer is closed on this GPU pool; this is the substitute.
Build first (here, no GPU needed):  python -m lrp_imagecaptioning_b200.build --debug
Then on the GPU box:                python tools/run_bounds_check.py > gpurun_out/bounds_check.log"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "lrp_imagecaptioning_b200", "liblrpcap_dbg.so")
assert os.path.exists(lib), "build the debug library first: python -m lrp_imagecaptioning_b200.build --debug"
env = dict(os.environ, LRPCAP_LIB=lib)
sel = ["tests/test_gpu_conv.py", "tests/test_gpu_encoder.py::test_relevance_matches_oracle_small",
       "tests/test_gpu_encoder.py::test_vgg19_relevance_matches_oracle", "tests/test_gpu_encoder.py::test_chunking_is_invisible",
       "tests/test_gpu_explainers.py", "tests/test_gpu_decoder.py"]
r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-x", "-k", "not 224"] + sel, cwd=ROOT, env=env, capture_output=True, text=True)
print(r.stdout[-3000:])
print(r.stderr[-2000:])
print("library:", lib, "exit code:", r.returncode)
sys.exit(r.returncode)
