"""The reference's own artefacts, prepared for the boundary tests.  Test infrastructure only.

Everything lands under baseline/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the built .so):

  baseline/_ref/cwipc/                      the UNMODIFIED reference Python package (python/cwipc), installed with pip
                                            (--no-index --no-deps --target) from a /tmp copy of /root/reference/python
  baseline/_ref/python/test_cwipc_util.py   the reference's acceptance tests, as they are (python/test_cwipc_util.py)
  baseline/_ref/tests/fixtures/input/pcl_frame1.ply
                                            stand-in for the fixture the reference tree lacks: a 160 000-point
                                            synthetic cloud in pcl::PLYWriter's layout (written by make_ply_fixtures' header code)
  baseline/_ref/apps_bin/<app>              the reference's C++ apps (apps/<app>/<app>.cpp) compiled where they lie
                                            against include/cwipc_util/api.h and linked with libcwipc_util_cuda

Nothing is copied into tracked paths.  On the GPU box /root/reference does not exist: the tests use what was prepared
here (by __graft_entry__.build() in the container) and skip when it is absent.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
REF_DIR = os.path.join(REPO, "baseline", "_ref")
APPS_BIN = os.path.join(REF_DIR, "apps_bin")
LIB_DIR = os.path.join(REPO, "cwipc_util_b200", "lib")
STUBS = os.path.join(REPO, "tests", "stubs")
# the apps of the hot path, the generator that feeds them, and the two PCL-free checks (one of them plain C: it pins the C side
# of the header); cwipc_ply2dump / cwipc_dump2ply / cwipc_pcl2dump include PCL headers themselves and cannot be built here
APPS = ["cwipc_generate", "cwipc_downsample", "cwipc_remove_outliers", "cwipc_tilefilter", "cwipc_util_install_check", "cwipc_ply2dump_c"]


def reference_present() -> bool:
    return os.path.isdir(os.path.join(REFERENCE, "apps")) and os.path.isdir(os.path.join(REFERENCE, "python", "cwipc"))


def python_installed() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "cwipc", "util.py")) and os.path.exists(os.path.join(REF_DIR, "python", "test_cwipc_util.py"))


def apps_built() -> bool:
    return all(os.path.exists(os.path.join(APPS_BIN, a)) for a in APPS)


def install_python(force: bool = False) -> None:
    if python_installed() and not force:
        return
    tmp = "/tmp/cwipc_reference_python_copy"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.copytree(os.path.join(REFERENCE, "python"), tmp)   # pip wants to write egg-info next to setup.py; the tree is read-only
    os.makedirs(REF_DIR, exist_ok=True)
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--upgrade", "--find-links", "/opt/wheelhouse",
                        "--target", REF_DIR, tmp], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("pip install of the reference python package failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    os.makedirs(os.path.join(REF_DIR, "python"), exist_ok=True)
    shutil.copyfile(os.path.join(REFERENCE, "python", "test_cwipc_util.py"), os.path.join(REF_DIR, "python", "test_cwipc_util.py"))
    shutil.rmtree(tmp, ignore_errors=True)


def write_fixture() -> str:
    """pcl_frame1.ply stand-in (the reference's own fixture is not in its tree): PCL layout, ASCII, 160 000 points."""
    sys.path.insert(0, os.path.join(REPO, "tests", "golden"))
    sys.path.insert(0, REPO)
    import make_ply_fixtures
    from cwipc_util_b200 import synthetic
    d = os.path.join(REF_DIR, "tests", "fixtures", "input")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "pcl_frame1.ply")
    if os.path.exists(path):
        return path
    pts = synthetic.synthetic_cloud(160000)
    with open(path + ".tmp", "w") as f:
        f.write(make_ply_fixtures.pcl_header("ascii", len(pts)))
        for p in pts:
            f.write(f"{float(p['x'])!r} {float(p['y'])!r} {float(p['z'])!r} {p['r']} {p['g']} {p['b']} {p['tile']}\n")
        f.write("0 0 0 1 0 0 0 1 0 0 0 1 0 0 0 0 0 %d 1 0 0\n" % len(pts))
    os.replace(path + ".tmp", path)
    return path


def build_apps(force: bool = False) -> None:
    """g++ on the reference's app sources where they lie; the only include path is OUR include/ (whose
    cwipc_util/api.h forwards to cwipc_util_cuda.h), the only library OUR libcwipc_util_cuda."""
    os.makedirs(APPS_BIN, exist_ok=True)
    for app in APPS:
        out = os.path.join(APPS_BIN, app)
        src = os.path.join(REFERENCE, "apps", app, app + ".cpp")
        cc = ["g++", "-std=c++17"]
        if not os.path.exists(src):
            src, cc = os.path.join(REFERENCE, "apps", app, app + ".c"), ["gcc", "-std=c11"]
        if os.path.exists(out) and not force and os.path.getmtime(out) >= os.path.getmtime(os.path.join(REPO, "include", "cwipc_util_cuda.h")):
            continue
        cmd = [*cc, "-O1", "-I", os.path.join(REPO, "include"), src, "-o", out + ".tmp", "-L", LIB_DIR, "-lcwipc_util_cuda",
               "-Wl,-rpath,$ORIGIN/../../../cwipc_util_b200/lib"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference app {app} does not compile/link against include/ + libcwipc_util_cuda:\n{r.stderr[-3000:]}")
        os.replace(out + ".tmp", out)


def prepare() -> None:
    """Called by __graft_entry__.build() when the reference tree is present."""
    if not reference_present():
        return
    install_python()
    write_fixture()
    build_apps()


def reference_env() -> dict:
    """Environment in which the unmodified reference package finds our library BY NAME: python/cwipc/util.py:149-161
    resolves find_library('cwipc_util') through LD_LIBRARY_PATH, where cwipc_util_b200/lib/libcwipc_util.so sits."""
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = LIB_DIR + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    env["PYTHONPATH"] = os.pathsep.join([STUBS, REF_DIR, env.get("PYTHONPATH", "")])
    return env


if __name__ == "__main__":
    prepare()
    print("python:", python_installed(), "apps:", apps_built())
