"""Build libsnb_b200 of another git revision next to the working tree's, for A/B timing on the same GPU box:
    python tools/build_rev.py HEAD            # -> tools/ab/libsnb_b200_HEAD.so   (git-ignored, travels with gpurun)
    SNB_B200_LIB=tools/ab/libsnb_b200_HEAD.so python tools/layer_times.py 13
    python tools/build_rev.py --profile       # working tree with -DSNB_CONV_PROFILE -> tools/ab/libsnb_b200_profile.so
Only meaningful while both revisions export the same C ABI."""
import importlib.util
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "segmentation-networks-benchmark_b200"
rev = sys.argv[1] if len(sys.argv) > 1 else "HEAD"
profile = rev == "--profile"
spec = importlib.util.spec_from_file_location("snb_build", os.path.join(ROOT, PKG, "build.py"))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
out_dir = os.path.join(ROOT, "tools", "ab")
os.makedirs(out_dir, exist_ok=True)
with tempfile.TemporaryDirectory() as tmp:
    if profile:
        subprocess.run(["cp", "-r", "--parents", PKG + "/csrc", "include", tmp], cwd=ROOT, check=True)
        rev = "profile"
    else:
        tar = subprocess.run(["git", "-C", ROOT, "archive", rev, PKG + "/csrc", "include"], check=True, stdout=subprocess.PIPE).stdout
        subprocess.run(["tar", "-x", "-C", tmp], input=tar, check=True)
    csrc = os.path.join(tmp, PKG, "csrc")
    flags = [f.replace(os.path.join(ROOT, "include"), os.path.join(tmp, "include")) for f in b.NVCC_FLAGS if f != "--use_fast_math=false"]
    if profile:
        flags.append("-DSNB_CONV_PROFILE")
        flags += os.environ.get("SNB_EXTRA_NVCC_FLAGS", "").split()      # timing experiments (e.g. -DSNB_TS_EXP=1)
        if os.environ.get("SNB_PROFILE_TAG"):
            rev = "profile_" + os.environ["SNB_PROFILE_TAG"]
    objs, procs = [], []
    for src in sorted(f for f in os.listdir(csrc) if f.endswith(".cu")):
        obj = os.path.join(tmp, src[:-3] + ".o")
        objs.append(obj)
        procs.append(subprocess.Popen(["/usr/local/cuda/bin/nvcc"] + flags + ["-c", os.path.join(csrc, src), "-o", obj]))
    for p in procs:
        if p.wait() != 0:
            sys.exit("nvcc failed")
    lib = os.path.join(out_dir, "libsnb_b200_%s.so" % rev.replace("/", "_"))
    subprocess.run(["/usr/local/cuda/bin/nvcc", "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], check=True)
    print(lib)
