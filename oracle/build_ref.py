"""Build oracle/_ref/libmct_ref.so = the UNMODIFIED reference sources + header shims + ref_harness.cpp.

TEST INFRASTRUCTURE ONLY.  The reference's own build system (obj/Makefile: Intel IPP, OpenCV 2.3.1, Boost) cannot
run in this image; this recipe compiles the reference translation units WHERE THEY LIE under /root/reference/src with
g++ against oracle/ref_shim/ (ipp.h, cv.h / cxcore.h / highgui.h / ml.h, boost/shared_ptr.hpp).  No reference source
is copied into the repository; only object files and the shared library are written, into oracle/_ref/ (git-ignored,
NOT gpurun-ignored: the library travels to the GPU box like the project's own .so files).

-ffp-contract=off: no FMA contraction, as in the reference's x86-64 SSE2 builds (obj/Makefile has no -march).
-O2: the shipped Makefile passes no -O flag; IEEE f32/f64 arithmetic does not depend on it under SSE2.
OpenMP stays off (obj/Makefile has no -fopenmp): the pragmas are ignored, as in the shipped build.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("MCT_REFERENCE_SRC", "/root/reference/src")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libmct_ref.so")                  # std::sort ties in index order (see ref_shim/boost/shared_ptr.hpp)
LIB_NATIVE = os.path.join(OUT, "libmct_ref_native.so")    # libstdc++ introsort tie order
# every reference translation unit except the command-line main
TUS = ["Public", "Config", "Matrix", "Sample", "SampleSet", "FeatureParameters", "HaarFeature", "HaarFeatureVector",
       "MultiDimensionalColorHistogram", "MultiDimensionalColorHistogramFeatureVector", "CultureColorHistogram",
       "CultureColorHistogramFeatureVector", "HaarAndColorHistogramFeatureVector", "WeakClassifierBase",
       "OnlineStumpsWeakClassifier", "WeightedStumpsWeakClassifier", "PerceptronWeakClassifier", "StrongClassifierBase",
       "StrongClassifierFactory", "MILBoostClassifier", "MILAnyBoostClassifier", "MILEnsembleClassifier", "AdaBoostClassifier",
       "ParticleFilter", "TrackerParameters", "Tracker", "SimpleTracker", "ParticleFilterTracker",
       "AppearanceBasedInformationFuser", "GeometryBasedInformationFuser", "Object", "Camera", "CameraNetwork"]
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = ["-std=gnu++11", "-O2", "-ffp-contract=off", "-fPIC", "-w", "-fpermissive", "-I" + os.path.join(HERE, "ref_shim")]


def available():
    return os.path.isdir(REF_SRC)


def _own_sources():
    return [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "ref_shim", "mct_cvshim.cpp"),
            os.path.join(HERE, "ref_shim", "mct_cvshim.h"), os.path.join(HERE, "ref_shim", "ipp.h"),
            os.path.join(HERE, "ref_shim", "boost", "shared_ptr.hpp"), os.path.abspath(__file__)]


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(s) > t for s in _own_sources())


def build(force=False, verbose=False, native=False):
    """Returns the library path, or None when neither the reference sources nor a prebuilt library are present."""
    lib = LIB_NATIVE if native else LIB
    if not available():
        return lib if os.path.exists(lib) else None
    if not force and not needs_build(lib):
        return lib
    odir = os.path.join(OUT, "obj_native" if native else "obj")
    os.makedirs(odir, exist_ok=True)
    ties = [] if native else ["-DMCT_REF_STABLE_TIES"]
    jobs = [(os.path.join(REF_SRC, t + ".cpp"), os.path.join(odir, t + ".o"), ties) for t in TUS]
    jobs.append((os.path.join(HERE, "ref_shim", "mct_cvshim.cpp"), os.path.join(odir, "mct_cvshim.o"), []))
    jobs.append((os.path.join(HERE, "ref_harness.cpp"), os.path.join(odir, "ref_harness.o"), ["-I" + REF_SRC] + ties))

    def cc(job):
        src, obj, extra = job
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max([os.path.getmtime(src)] + [os.path.getmtime(s) for s in _own_sources()[2:]]):
            return None
        r = subprocess.run([CXX] + FLAGS + extra + ["-c", src, "-o", obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed on %s\n%s" % (src, r.stdout[-4000:]))
        return src

    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 4)) as ex:
        done = [d for d in ex.map(cc, jobs) if d]
    tmp = lib + ".tmp.%d" % os.getpid()
    r = subprocess.run([CXX, "-shared", "-o", tmp] + [j[1] for j in jobs] + ["-lm"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed\n" + r.stdout[-4000:])
    os.replace(tmp, lib)
    if verbose:
        print("compiled %d translation units -> %s" % (len(done), lib))
    return lib


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose=True))
    print(build(force="-f" in sys.argv, verbose=True, native=True))
