"""TEST INFRASTRUCTURE: builds tests/emu/libpatch_emu.so = the product's device header compiled by g++ with the SIMT
emulation (simt_emu.h)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libpatch_emu.so")
DEPS = [os.path.join(HERE, "patch_emu.cc"), os.path.join(HERE, "simt_emu.h"),
        os.path.join(os.path.dirname(os.path.dirname(HERE)), "mve_b200", "csrc", "patch_opt.cuh"),
        os.path.join(os.path.dirname(os.path.dirname(HERE)), "mve_b200", "csrc", "patch_warp.cuh"),
        os.path.join(os.path.dirname(os.path.dirname(HERE)), "mve_b200", "csrc", "patch_thread.cuh")]


def build(force=False):
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I" + HERE,
                           os.path.join(HERE, "patch_emu.cc"), "-o", LIB])
    return LIB
