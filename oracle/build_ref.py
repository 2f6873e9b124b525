#!/usr/bin/env python3
"""Build the UNMODIFIED reference (AMIE) solver objects out-of-tree into oracle/_ref/.

Test infrastructure only (see oracle/README.md).  Nothing is copied from /root/reference:
the sources are compiled where they lie; outputs go to oracle/_ref/ (git-ignored).

Recipe:
  * source list = ALL_SRC of <ref>/CMakeLists.txt (parsed at build time; cmake itself is
    not run -- cmake >= 4 rejects that file);
  * flags = the reference's release flags (CMakeLists.txt:329-362) minus the SSE defines
    (-DHAVE_SSE3 branches do not compile);
  * geometry/space_time_geometry_2D.cpp is left out of the archive because
    physics/dual_behaviour.cpp:10 #includes it;
  * the archive is linked with oracle/ref_harness.cpp (ours) into
    oracle/_ref/libamie_ref_oracle.so, a C-ABI around Amie::ConjugateGradient,
    Amie::BiConjugateGradientStabilized, assign(y, A*x[-b]) and inverseDiagonal().
"""
import os, re, subprocess, sys, concurrent.futures as cf

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("AMIE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "obj")
# the image exports CXX=/opt/gcc/bin/g++ (a wrapper without libgomp.spec): use the system driver
CXX = os.environ.get("AMIE_CXX", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")
FLAGS = ["-std=c++17", "-fext-numeric-literals", "-O3", "-funroll-loops", "-ftree-vectorize",
         "-fopenmp", "-DHAVE_OPENMP", "-DNDEBUG", "-fPIC", "-w"]
SKIP = {"geometry/space_time_geometry_2D.cpp"}


def all_src():
    txt = open(os.path.join(REF, "CMakeLists.txt")).read()
    m = re.search(r"set\(ALL_SRC(.*?)\n\)", txt, re.S)
    seen, out = set(), []
    for tok in m.group(1).split():
        if tok.startswith("#") or not tok.endswith(".cpp"):
            continue
        if tok in seen or tok in SKIP:
            continue
        seen.add(tok)
        out.append(tok)
    return out


def compile_one(src):
    obj = os.path.join(OBJ, src.replace("/", "__")[:-4] + ".o")
    s = os.path.join(REF, src)
    if os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(s):
        return obj, 0, ""
    p = subprocess.run([CXX, *FLAGS, "-c", s, "-o", obj], capture_output=True, text=True)
    return obj, p.returncode, p.stderr[-2000:]


def main():
    if not os.path.isdir(REF):
        print(f"[build_ref] {REF} absent: keeping prebuilt oracle/_ref as is")
        return 0
    os.makedirs(OBJ, exist_ok=True)
    srcs = all_src()
    jobs = int(os.environ.get("JOBS", os.cpu_count() or 4))
    objs, bad = [], []
    with cf.ThreadPoolExecutor(jobs) as ex:
        for obj, rc, err in ex.map(compile_one, srcs):
            (objs if rc == 0 else bad).append(obj)
            if rc:
                print("[build_ref] FAILED", obj, err, file=sys.stderr)
    if bad:
        return 1
    lib = os.path.join(OUT, "libAmie.a")
    if os.path.exists(lib):
        os.remove(lib)
    subprocess.check_call(["ar", "rcs", lib, *objs])
    so = os.path.join(OUT, "libamie_ref_oracle.so")
    subprocess.check_call([CXX, *FLAGS, "-shared", "-I" + REF, os.path.join(HERE, "ref_harness.cpp"),
                           lib, "-Wl,--no-undefined", "-lm", "-o", so])
    print("[build_ref] built", so, f"({len(objs)} reference objects)")

    # ---- end-to-end pair: the SAME FeatureTree driver linked with the reference solvers / with the drop-in TUs
    harness = os.path.join(HERE, "e2e_harness.cpp")
    e2e_ref = os.path.join(OUT, "amie_e2e_ref")
    subprocess.check_call([CXX, *FLAGS, "-I" + REF, harness, lib, "-lm", "-o", e2e_ref])
    print("[build_ref] built", e2e_ref)
    pkg = os.path.join(os.path.dirname(HERE), "xfem-amie_b200")
    b200 = os.path.join(pkg, "libamie_b200.so")
    if os.path.exists(b200):
        shim_objs = []
        for src in ("amie_b200_shim.cpp", "conjugategradient_b200.cpp", "biconjugategradientstabilized_b200.cpp"):
            o = os.path.join(OBJ, "shim__" + src[:-4] + ".o")
            subprocess.check_call([CXX, *FLAGS, "-I" + REF, "-c", os.path.join(pkg, "host", "shim", src), "-o", o])
            shim_objs.append(o)
        e2e_b200 = os.path.join(OUT, "amie_e2e_b200")
        # the shim objects come first: the archive's conjugategradient.o / biconjugategradientstabilized.o are then
        # never pulled in (every symbol they define is already defined)
        subprocess.check_call([CXX, *FLAGS, "-DAMIE_B200_E2E", "-I" + REF, "-I" + os.path.join(pkg, "host", "shim"), harness,
                               *shim_objs, lib, "-L" + pkg, "-lamie_b200",
                               "-Wl,-rpath,$ORIGIN/../../xfem-amie_b200", "-lm", "-o", e2e_b200])
        print("[build_ref] built", e2e_b200)
    return 0


if __name__ == "__main__":
    sys.exit(main())
