"""Diagnostic (build container only -- it reads /root/reference): run the reference's OWN test files, unmodified, with the
names ``CurvlinopsGGN`` / ``CurvlinopsEF`` rebound to ``B200GGN`` / ``B200EF`` everywhere the reference looks them up (the
default backend of ``BaseLaplace``, the imports of its test modules), on the CPU emulation of the kernels.

    python tests/diagnostics/run_reference_suite.py [pytest args ...]      # default: the files on the hot path

Cases parametrised over BackPACK / ASDL / asdfghjkl backends hit the inert placeholders of those absent libraries and fail
or error by construction; the summary separates them (by test id) from the cases that exercise our classes.  The committed,
self-contained counterpart of this run is tests/test_reference_*_cases_cpu.py."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("LAPLACE_REFERENCE_ROOT", "/root/reference")
FILES = ["test_baselaplace.py", "test_lllaplace.py", "test_matrix.py", "test_curv_backends_curvlinops.py", "test_serialization.py",
         "test_subset_params.py", "test_subnetlaplace.py", "test_laplace.py", "test_utils.py"]
FOREIGN = re.compile(r"Asdl|BackPack|Asdfghjkl|asdl|backpack|Hessian|LowRank|Functional|large|SWAG", re.I)


class Plugin:
    """Installs the emulated kernels and the rebinding before the reference's test modules are imported."""

    def __init__(self):
        self.outcomes = {}

    def pytest_sessionstart(self, session):
        import torch

        sys.path.insert(0, ROOT)
        import tests.cpu_kernels as ck           # our tests package first: the reference ships a ``tests`` package too
        from oracle import ref_shim

        ref_shim.install()

        class Patch:
            def setattr(self, obj, name, value):
                setattr(obj, name, value)

        ck.install(Patch())
        import laplace
        import laplace.baselaplace
        import laplace.curvature
        import laplace.curvature.curvlinops as cl
        import laplace.lllaplace
        import laplace.subnetlaplace

        from laplace_b200 import B200EF, B200GGN

        for mod in (cl, laplace.curvature, laplace.baselaplace, laplace.lllaplace, laplace.subnetlaplace, laplace):
            for name, cls in (("CurvlinopsGGN", B200GGN), ("CurvlinopsEF", B200EF)):
                if hasattr(mod, name):
                    setattr(mod, name, cls)
        # the reference's tests import helpers from ITS tests package
        for k in [k for k in sys.modules if k == "tests" or k.startswith("tests.")]:
            del sys.modules[k]
        sys.path.insert(0, REF)
        torch.set_default_dtype(torch.double)     # what its test modules set at import time

    def pytest_runtest_logreport(self, report):
        if report.when == "call" or (report.when == "setup" and report.outcome != "passed"):
            self.outcomes[report.nodeid] = report.outcome


def run(args):
    import pytest

    plugin = Plugin()
    os.chdir(REF)
    pytest.main(["-q", "-p", "no:cacheprovider", "--no-header", "-rN", "--tb=no", "-W", "ignore", "--rootdir", REF, "-c",
                 os.devnull] + args, plugins=[plugin])
    return plugin.outcomes


def tally(d):
    out = {}
    for v in d.values():
        out[v] = out.get(v, 0) + 1
    return out


def main():
    import json
    import subprocess

    if sys.argv[1:2] == ["--one"]:                # child: one pytest session, outcomes as JSON on the last line
        print("OUTCOMES " + json.dumps(run(sys.argv[2:])))
        return 0
    if sys.argv[1:]:
        outcomes = run(sys.argv[1:])
    else:                                         # one session per file: the reference's modules set process-global state
        outcomes = {}                             # (default dtype, seeds) at import time and do not restore it
        for f in FILES:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", os.path.join(REF, "tests", f)],
                               capture_output=True, text=True)
            line = [ln for ln in r.stdout.splitlines() if ln.startswith("OUTCOMES ")]
            got = json.loads(line[-1][9:]) if line else {}
            print(f"{f:36s} {tally(got)}", flush=True)
            outcomes.update(got)
    ours = {k: v for k, v in outcomes.items() if not FOREIGN.search(k)}
    foreign = {k: v for k, v in outcomes.items() if FOREIGN.search(k)}
    print("\ncases that exercise the B200 classes:", tally(ours))
    print("cases parametrised over absent third-party backends / out-of-scope classes:", tally(foreign))
    bad = sorted(k for k, v in ours.items() if v == "failed")
    for k in bad[:60]:
        print("  FAILED", k.replace(REF + "/", ""))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
