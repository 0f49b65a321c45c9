"""Runs the REFERENCE'S OWN unit tests (/root/reference/tests/test_*.py, unmodified) against the reference's own code
with ``tensorflow`` served by tests/golden/tf_numpy_shim.py.  Purpose: to show that the shim behaves like TensorFlow for
this code base -- the reference's literal known answers (exact alpha / beta tables, losses 0 / 100.0 / 1e10 / +inf, exact
gradients, tools.py examples) come out right when its code runs on the shim -- so that the fixtures generated through
the same shim (make_reference_golden.py) can be trusted as outputs of the reference.

    python tests/golden/run_reference_tests.py [--float32]        (needs /root/reference; build container only)

The run is in float64 like the fixture generation (``tf.float32`` names float64 in the shim).  With --float32 two tests
miss their 8-decimal bar by 2e-8: they subtract a numpy float64 constant from a float32 loss, which TensorFlow would
first convert to float32 and numpy promotes to float64 -- the one place where the shim's float32 mode is not TensorFlow's.

Tests that need TensorFlow machinery the shim does not have -- tf.GradientTape (autodiff), tf.nn.ctc_loss, jacobians --
are reported as "needs TensorFlow" and not counted; every other test must pass.  Exit status 1 if a runnable test fails.
Always run as a separate process: it puts /root/reference first on sys.path (the reference's test modules import each
other as ``tests.*``) and registers the shim as ``tensorflow``."""
from __future__ import annotations

import os
import sys
import unittest
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
MODULES = ["tests.test_tools", "tests.test_classic_ctc_loss", "tests.test_simplified_ctc_loss", "tests.test_hessian"]
NEEDS_TF = ("GradientTape", "ctc_loss", "jacobian", "batch_jacobian", "map_fn", "'nn'")


def main():
    repo_root = os.path.dirname(os.path.dirname(HERE))
    sys.path[:] = [HERE, REFERENCE] + [p for p in sys.path if os.path.abspath(p or ".") not in (repo_root, HERE, REFERENCE)]
    os.chdir(REFERENCE)
    import numpy as np
    import tf_numpy_shim
    tf_numpy_shim.install(np.float32 if "--float32" in sys.argv else np.float64)
    warnings.simplefilter("ignore")
    passed, failed, needs_tf = [], [], []
    for mod in MODULES:
        try:
            suite = unittest.defaultTestLoader.loadTestsFromName(mod)
        except Exception as e:  # noqa: BLE001
            needs_tf.append((mod, f"import: {type(e).__name__}: {e}"))
            continue
        stack = [suite]
        while stack:
            item = stack.pop()
            if isinstance(item, unittest.TestSuite):
                stack.extend(item)
                continue
            if item.__class__.__name__ == "_FailedTest":
                needs_tf.append((item.id(), "module import failed"))
                continue
            result = unittest.TestResult()
            item.run(result)
            if result.wasSuccessful() and not result.skipped:
                passed.append(item.id())
                continue
            text = "".join(t for _, t in result.errors + result.failures)
            last = text.strip().splitlines()[-1] if text.strip() else "skipped"
            if result.errors and any(k in text for k in NEEDS_TF):
                needs_tf.append((item.id(), last))
            else:
                failed.append((item.id(), last))
    for name in sorted(passed):
        print(f"PASS      {name}")
    for name, why in sorted(needs_tf):
        print(f"NEEDS-TF  {name}    ({why[:110]})")
    for name, why in sorted(failed):
        print(f"FAIL      {name}    ({why[:160]})")
    print(f"reference tests under the numpy shim: {len(passed)} passed, {len(failed)} failed, "
          f"{len(needs_tf)} need TensorFlow machinery the shim does not provide")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
