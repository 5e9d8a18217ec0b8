"""Build the product sources with extra -D flags into tools/_build/lib_<tag>/libb200rl.so (A/B timing of compile-time
variants through B200RL_LIB=<path>; dev tool).   python tools/build_variant.py ni4 -DB200RL_BWD_NI=4"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-distributed-reinforcement-learning_b200"))
from b200rl import build as B

tag, defs = sys.argv[1], sys.argv[2:]
out = os.path.join(ROOT, "tools", "_build", "lib_" + tag)
os.makedirs(out, exist_ok=True)
objs = []


def cc(src):
    obj = os.path.join(out, os.path.basename(src)[:-3] + ".o")
    r = subprocess.run([B.NVCC] + B.FLAGS + defs + ["-c", src, "-o", obj], capture_output=True, text=True)
    if r.returncode:
        raise SystemExit(r.stderr)
    return obj


with ThreadPoolExecutor(8) as ex:
    objs = list(ex.map(cc, B._sources()))
lib = os.path.join(out, "libb200rl.so")
subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + ["-lcudart"], check=True)
for o in objs:
    os.remove(o)
print(lib)
