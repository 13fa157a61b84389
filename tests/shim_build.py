"""Builds libjoltc_gpx.so (the joltc subset over libgpx) and the headless engine stand-in that exercises it."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "c-game-engine_b200")
SHIM_LIB = os.path.join(PKG, "libjoltc_gpx.so")
DRIVER_SRC = os.path.join(ROOT, "tests", "shim", "engine_calls.c")
DRIVER = os.path.join(ROOT, "tests", "shim", "_build", "engine_calls")


def build_shim() -> str:
    r = subprocess.run(["make", "-C", os.path.join(PKG, "shim")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libjoltc_gpx.so failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return SHIM_LIB


def build_driver() -> str:
    """gcc, C11, warnings as errors: the engine-style source must compile against include/joltc/ unchanged."""
    build_shim()
    os.makedirs(os.path.dirname(DRIVER), exist_ok=True)
    if os.path.exists(DRIVER) and os.path.getmtime(DRIVER) >= max(
            os.path.getmtime(DRIVER_SRC), os.path.getmtime(SHIM_LIB), os.path.getmtime(os.path.join(ROOT, "include", "joltc_gpx.h"))):
        return DRIVER
    cmd = ["gcc", "-std=gnu11", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), DRIVER_SRC, "-o", DRIVER,
           "-L", PKG, "-ljoltc_gpx", "-lgpx", "-lm", f"-Wl,-rpath,{PKG}", "-Wl,-rpath,$ORIGIN/../../../c-game-engine_b200"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building the engine stand-in failed:\n" + r.stderr[-4000:])
    return DRIVER
