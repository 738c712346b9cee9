"""Build csrc/libtrajopt_b200.so with nvcc for sm_100a (cross-compiles without a GPU).

One translation unit per kernel instantiation (model × integrator × ALTRO transform) so they
compile in parallel; `registry.cu` is generated from the same list.  Flags: -fmad=false keeps the
arithmetic contract (FMAs only where fma() is written), -lineinfo lets ncu map SASS to source.
"""
import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "_gen")
LIB = os.path.join(CSRC, "libtrajopt_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]

# (model, integrator, infeasible, min_time, partials per lane)
INSTANCES = [
    # rk3: every model, plus the ALTRO transforms the problem zoo uses (infeasible start, minimum time)
    (0, 0, 0, 0, 3),
    (1, 0, 0, 0, 3), (1, 0, 0, 1, 4),
    (2, 0, 0, 0, 5), (2, 0, 1, 0, 5), (2, 0, 0, 1, 6), (2, 0, 1, 1, 6),   # car: + infeasible start AND minimum time in one solve
    (1, 0, 1, 0, 3), (1, 0, 1, 1, 4),
    (3, 0, 0, 0, 5),
    (4, 0, 0, 0, 2),
    (5, 0, 0, 0, 5), (5, 0, 0, 1, 6),
    (6, 0, 0, 0, 6), (6, 0, 0, 1, 7),
    # rk4 and midpoint (src/integration.jl:115-125, 26-33): every model
    (0, 1, 0, 0, 3), (0, 2, 0, 0, 3),
    (1, 1, 0, 0, 3), (1, 2, 0, 0, 3),
    (2, 1, 0, 0, 5), (2, 2, 0, 0, 5),
    (3, 1, 0, 0, 5), (3, 2, 0, 0, 5),
    (4, 1, 0, 0, 2), (4, 2, 0, 0, 2),
    (5, 1, 0, 0, 5), (5, 2, 0, 0, 5),
    (6, 1, 0, 0, 6), (6, 2, 0, 0, 6),
]


def _name(i):
    return "m%d_i%d_f%d_t%d" % i[:4]


def _write(path, text):
    if not os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)
    return path


def generate():
    os.makedirs(GEN, exist_ok=True)
    files = []
    for inst in INSTANCES:
        mo, ig, inf, mt, pc = inst
        src = ('#include "../lockstep.cuh"\nnamespace tob {\nKernelInfo tob_info_%s() { return make_info<Cfg<%d, %d, %s, %s, %d>>(); }\n}\n'
               % (_name(inst), mo, ig, "true" if inf else "false", "true" if mt else "false", pc))
        files.append(_write(os.path.join(GEN, "inst_%s.cu" % _name(inst)), src))
    reg = ['#include <vector>\n#include "../engine_host.h"\nnamespace tob {']
    reg += ["KernelInfo tob_info_%s();" % _name(i) for i in INSTANCES]
    reg.append("const KernelInfo* find_kernel(int model, int integ, int inf, int mt) {")
    reg.append("    static const std::vector<KernelInfo> all = {%s};" % ", ".join("tob_info_%s()" % _name(i) for i in INSTANCES))
    reg.append("    for (const auto& k : all) if (k.model == model && k.integ == integ && k.inf == inf && k.mt == mt) return &k;")
    reg.append("    return nullptr;\n}")
    reg.append("void model_dims(int model, int* n, int* m) {")
    reg.append("    static const int d[7][2] = {{2, 1}, {2, 1}, {3, 2}, {4, 1}, {13, 4}, {4, 1}, {4, 2}};")
    reg.append("    *n = d[model][0]; *m = d[model][1];\n}\n}")
    files.append(_write(os.path.join(GEN, "registry.cu"), "\n".join(reg) + "\n"))
    files.append(os.path.join(CSRC, "capi.cu"))
    files.append(os.path.join(CSRC, "peak.cu"))
    return files


KERNEL_DEPS = ("engine.cuh", "lockstep.cuh", "resident.cuh", "pn.cuh", "sqrt_bp.cuh", "engine_host.h", "models.cuh", "../build.py", "../../include/trajopt_b200.h")
HOST_DEPS = ("engine_host.h", "../build.py", "../../include/trajopt_b200.h")


def _hash(src):
    """hash of one translation unit and the headers it includes (instances: every kernel header; capi/peak/registry: host headers)"""
    h = hashlib.sha1()
    deps = KERNEL_DEPS if os.path.basename(src).startswith("inst_") else HOST_DEPS
    for f in (src,) + tuple(os.path.join(CSRC, d) for d in deps):
        h.update(open(f, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(GEN, os.path.basename(src).replace(".cu", ".o"))
    stamp, hsh = obj + ".stamp", _hash(src)
    log = obj + ".log"
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == hsh:
        return src, obj, 0, open(log).read() if os.path.exists(log) else "", False
    cmd = [NVCC] + ARCH + FLAGS + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode == 0:
        open(stamp, "w").write(hsh)
        open(log, "w").write(r.stdout + r.stderr)
    return src, obj, r.returncode, r.stdout + r.stderr, True


def build(force=False, verbose=False, jobs=None):
    files = generate()
    if force:
        for f in os.listdir(GEN):
            if f.endswith(".stamp"):
                os.remove(os.path.join(GEN, f))
    jobs = jobs or max(1, (os.cpu_count() or 4))
    objs, logs, rebuilt = [], [], False
    with cf.ThreadPoolExecutor(jobs) as ex:
        for src, obj, rc, out, did in ex.map(_compile, files):
            logs.append("== %s\n%s" % (os.path.basename(src), out))
            if rc != 0:
                sys.stderr.write(out)
                raise RuntimeError("nvcc failed on %s" % src)
            objs.append(obj)
            rebuilt = rebuilt or did
    open(os.path.join(GEN, "ptxas.log"), "w").write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
