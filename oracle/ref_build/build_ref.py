"""Compile the reference's own cpp_wrappers sources, where they lie under /root/reference, into
oracle/_ref/libbxref.so (git-ignored; travels to the GPU box with the snapshot).  g++ on the few source files
directly -- the reference's setup.py / TBB / Eigen are not used: oracle/ref_build/shim/ provides the handful of
tbb:: and Eigen:: names the sources touch.  No reference source is copied into this repository."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/cpp_wrappers"
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
OUT = os.path.join(OUT_DIR, "libbxref.so")


def build(force=False):
    if not os.path.isdir(REF):
        raise RuntimeError("/root/reference is not present (GPU box): the prebuilt oracle/_ref/libbxref.so is used")
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, "ref_api.cpp"), os.path.join(REF, "cpp_neighbors/neighbors/neighbors.cpp"),
            os.path.join(REF, "cpp_subsampling/grid_subsampling/grid_subsampling.cpp"), os.path.join(REF, "cpp_utils/cloud/cloud.cpp")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) > os.path.getmtime(s) for s in srcs):
        return OUT
    cmd = ["g++", "-O3", "-std=c++17", "-fPIC", "-shared", "-pthread", "-w", "-I", os.path.join(HERE, "shim"), "-I", REF,
           "-o", OUT] + srcs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
