"""Build libmct_host.so: the C++ host mirror of the reference interfaces (g++), linked against libmct_b200.so."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libmct_host.so")
SRC = [os.path.join(HERE, "mct_host.cpp"), os.path.join(HERE, "mct_fuser.cpp"), os.path.join(HERE, "mct_host_c.cpp"),
       os.path.join(HERE, "mct_network.cpp")]


def build(force=False):
    from .. import build as core
    core.build()
    deps = SRC + [os.path.join(HERE, "mct_host.h"), os.path.join(HERE, "mct_fuser.h"), os.path.join(HERE, "mct_host_c.h"),
                  os.path.join(HERE, "mct_network.h"), os.path.join(PKG, "..", "include", "mct_b200.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    import fcntl
    with open(os.path.join(HERE, ".build.lock"), "w") as lk:
        fcntl.flock(lk, fcntl.LOCK_EX)           # one builder at a time (ranks of a torchrun job)
        try:
            if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
                return LIB
            gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
            tmp = LIB + ".tmp.%d" % os.getpid()
            cmd = [gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-pthread", "-o", tmp] + SRC + [
                "-L" + PKG, "-l:libmct_b200.so", "-Wl,-rpath,$ORIGIN/.."]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            if r.returncode != 0:
                sys.stderr.write(r.stdout)
                raise RuntimeError("g++ failed building libmct_host.so")
            os.replace(tmp, LIB)
            return LIB
        finally:
            fcntl.flock(lk, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force=True))
