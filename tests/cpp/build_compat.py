"""Build the two programs that prove the drop-in boundary into tests/cpp/_build/:

  hector_cli   the reference's CLI, src/main.cpp, compiled UNMODIFIED against include/compat
  rcpp_glue    the reference's R glue, src/rcpp_hector.cpp, compiled UNMODIFIED against
               include/compat and the stub tests/cpp/rcpp_stub/Rcpp.h, plus a C++ driver that calls
               it the way the R package does (tests/cpp/test_rcpp_glue.cpp)

Needs the reference sources ($HECTOR_REFERENCE, default /root/reference): they are compiled where
they lie, nothing is copied.  The binaries are git-ignored and travel to the GPU box like the
library does.  Called by __graft_entry__.build() and tests/test_compat.py."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build")
REF = os.environ.get("HECTOR_REFERENCE", "/root/reference")
LIBDIR = os.path.join(ROOT, "hector_b200")


def reference_present():
    return os.path.exists(os.path.join(REF, "src", "main.cpp"))


def build_xchg():
    """tests/cpp/test_xchg.cpp: the push exchange driven from C++ by two forked processes (needs
    only this repo and the CUDA runtime headers) -> path"""
    os.makedirs(OUT, exist_ok=True)
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = os.path.join(OUT, "xchg_cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I",
                    os.path.join(cuda, "include"), os.path.join(HERE, "test_xchg.cpp"), "-L", LIBDIR,
                    "-lhector_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart",
                    "-Wl,-rpath," + LIBDIR, "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe],
                   check=True, capture_output=True, text=True)
    return exe


def build():
    """-> {name: path}; raises CalledProcessError with the compiler's output on failure"""
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include", "compat"), "-I", os.path.join(ROOT, "include")]
    link = ["-L", LIBDIR, "-lhector_b200", "-Wl,-rpath," + LIBDIR]
    cli = os.path.join(OUT, "hector_cli")
    subprocess.run(["g++", "-std=c++17", "-O1"] + inc + [os.path.join(REF, "src", "main.cpp")] + link
                   + ["-o", cli], check=True, capture_output=True, text=True)
    glue = os.path.join(OUT, "rcpp_glue")
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", os.path.join(HERE, "rcpp_stub")] + inc
                   + [os.path.join(REF, "src", "rcpp_hector.cpp"), os.path.join(HERE, "test_rcpp_glue.cpp")]
                   + link + ["-o", glue], check=True, capture_output=True, text=True)
    return {"hector_cli": cli, "rcpp_glue": glue}


if __name__ == "__main__":
    print(build_xchg())
    print(build())
