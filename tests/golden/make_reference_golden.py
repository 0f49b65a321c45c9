"""Generates tests/golden/reference/reference_outputs.npz by running the REFERENCE'S OWN CODE (/root/reference/tf_seq2seq_losses,
imported unmodified) on small seeded inputs.  TensorFlow cannot be installed in this image, so ``tensorflow`` is served
by tests/golden/tf_numpy_shim.py, a numpy implementation of the TensorFlow ops the reference calls; the reference's
Python -- the label cleaning, the masks, the tf.while_loop recursions, the transition tables, the token scatter, the
Hessian assembly -- runs as written.

    python tests/golden/make_reference_golden.py            (needs /root/reference; run in the build container only)

The pass runs in float64 (``tf.float32`` names float64 in the shim, so the reference's dtype assertion and casts follow):
the reference's arithmetic in double precision, which the oracle is compared with at 1e-12.  (A float32 pass would not be
faithful: numpy promotes ``x + np.log(2.0)``, tools.py:70, to float64 where TensorFlow keeps float32.)  Stored per case:
the inputs, and from the data classes (base_loss.py:186-298,
classic_ctc_loss.py:152-165,310-462, simplified_ctc_loss.py:73-83,291-438) ``loss``, ``gradient``,
``logarithmic_logproba_gradient``, ``alpha``, ``beta`` and -- small cases -- ``hessian`` and ``gamma``, plus the loss the
public functions classic_ctc_loss / simplified_ctc_loss return for the logits.

The fixtures travel; this script and the reference do not have to: tests/test_reference_golden.py compares the oracle
(CPU) and the CUDA path (GPU) with the stored outputs, and re-runs this generator only where /root/reference exists."""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference"
OUT = os.path.join(HERE, "reference", "reference_outputs.npz")     # its own directory: tests/test_golden.py globs tests/golden/*.npz

# name: (B, T, V, Lw, blank, seed, with_second_order)
CASES = {
    "small_ragged": (4, 12, 6, 5, 0, 3, True),
    "repeats": (3, 7, 4, 4, 0, 5, True),
    "blank_mid": (3, 9, 5, 3, 2, 7, True),
    "blank_last_empty_label": (3, 8, 5, 3, 4, 11, True),
    "labels_wider_than_needed": (2, 10, 6, 6, 0, 13, True),
    "mid": (3, 40, 12, 10, 0, 17, False),
    "long": (2, 300, 40, 60, 0, 19, False),          # first-order outputs only, no state tensors (fixture size)
    # undefined input with definite arithmetic: a real label EQUAL to the blank (base_loss.py:328-344 gathers its
    # log-probability like any token's; classic_ctc_loss.py:647-654 overrides the blank column of the gradient)
    "label_equals_blank": (3, 9, 5, 4, 1, 23, True),
}


def make_inputs(name):
    B, T, V, Lw, blank, seed, _ = CASES[name]
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((B, T, V)).astype(np.float32) * 2.0
    others = [k for k in range(V) if k != blank]
    labels = rng.choice(others, size=(B, Lw)).astype(np.int32)
    label_length = rng.integers(1, Lw + 1, size=B).astype(np.int32)
    logit_length = rng.integers(T // 2, T + 1, size=B).astype(np.int32)
    if name == "small_ragged":
        logit_length[0] = T
        label_length[1], logit_length[1] = 5, 4              # more labels than frames: infeasible, loss = +inf
    if name == "repeats":
        labels[0] = [1, 1, 2, 2]                             # "aabb" needs 6 frames in the classic loss, 4 in the simplified one
        label_length[0], logit_length[0] = 4, 5
        labels[1, :3] = [3, 3, 3]
        label_length[1], logit_length[1] = 3, 7
    if name == "blank_last_empty_label":
        label_length[0] = 0                                  # empty label: every frame is blank
        logit_length[2] = 0                                  # no frames at all
        label_length[2] = 0
    if name == "label_equals_blank":
        labels[0, 1] = blank
        labels[1, 0] = blank
        labels[1, 2] = blank
        label_length[:] = [3, 4, 2]
        logit_length[:] = [9, 8, 6]
    if name == "long":
        label_length[:] = [60, 41]
        logit_length[:] = [300, 233]
    if name == "labels_wider_than_needed":
        label_length[:] = [3, 2]                             # labels.shape[1] = 6 > max(label_length) + 1
    return logits, labels, label_length, logit_length, blank


def run_reference(float_dtype, names=None):
    """{case: {key: array}} computed by the reference under the numpy shim in `float_dtype` arithmetic."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import tf_numpy_shim
    tf_numpy_shim.install(float_dtype)
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)       # appended: the reference's own `tests` package must not shadow this repo's
    from tf_seq2seq_losses.classic_ctc_loss import ClassicCtcLossData, classic_ctc_loss
    from tf_seq2seq_losses.simplified_ctc_loss import SimplifiedCtcLossData, simplified_ctc_loss
    from tf_seq2seq_losses.tools import logit_to_logproba
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")     # log(0), inf - inf inside tf.where branches that are not selected
        for name in (names or CASES):
            logits, labels, ll, tl, blank = make_inputs(name)
            x = logits.astype(float_dtype)
            logprobas = logit_to_logproba(logit=x, axis=2)
            res = {}
            for tag, cls, fn in (("classic", ClassicCtcLossData, classic_ctc_loss),
                                 ("simplified", SimplifiedCtcLossData, simplified_ctc_loss)):
                data = cls(labels=labels, logprobas=logprobas, label_length=ll, logit_length=tl, blank_index=blank)
                keys = ["loss", "gradient", "logarithmic_logproba_gradient"] + (["alpha", "beta"] if x.shape[1] <= 64 else [])
                if CASES[name][6]:
                    keys += ["hessian", "gamma"]
                for k in keys:
                    res[f"{tag}/{k}"] = np.asarray(getattr(data, k))
                res[f"{tag}/public_loss"] = np.asarray(fn(labels, x, ll, tl, blank))
            out[name] = res
    return out


def build():
    f64 = run_reference(np.float64)
    flat = {}
    for name in CASES:
        logits, labels, ll, tl, blank = make_inputs(name)
        flat[f"{name}/logits"] = logits
        flat[f"{name}/labels"] = labels
        flat[f"{name}/label_length"] = ll
        flat[f"{name}/logit_length"] = tl
        flat[f"{name}/blank"] = np.int32(blank)
        for k, v in f64[name].items():
            flat[f"{name}/f64/{k}"] = v
    return flat


def check_oracle_made_fixtures():
    """tests/golden/*.npz were written by the oracle (make_golden.py).  Here the reference computes the same quantities for
    their inputs; returns the list of (file, key, max abs difference) that exceed 1e-12."""
    import glob
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    import tf_numpy_shim
    tf_numpy_shim.install(np.float64)
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)
    from tf_seq2seq_losses.classic_ctc_loss import ClassicCtcLossData
    from tf_seq2seq_losses.simplified_ctc_loss import SimplifiedCtcLossData
    from tf_seq2seq_losses.tools import logit_to_logproba
    bad, n = [], 0
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for path in sorted(glob.glob(os.path.join(HERE, "*.npz"))):
            z = np.load(path)
            logprobas = logit_to_logproba(logit=z["logits"].astype(np.float64), axis=2)
            for tag, cls in (("classic", ClassicCtcLossData), ("simplified", SimplifiedCtcLossData)):
                data = cls(labels=z["labels"], logprobas=logprobas, label_length=z["label_length"],
                           logit_length=z["logit_length"], blank_index=int(z["blank"]))
                for key in ("loss", "gradient", "alpha", "beta", "hessian"):
                    if f"{tag}_{key}" not in z.files:
                        continue
                    got, want = np.asarray(getattr(data, key)), z[f"{tag}_{key}"]
                    fin = np.isfinite(want)
                    diff = float(np.max(np.abs(got[fin] - want[fin]))) if fin.any() else 0.0
                    n += 1
                    if got.shape != want.shape or not np.array_equal(got[~fin], want[~fin]) or diff > 1e-12:
                        bad.append((os.path.basename(path), f"{tag}_{key}", diff))
    return n, bad


def fuzz(trials, seed=0):
    """Random small problems (every blank position, repeated tokens, ragged / infeasible / empty lengths): the oracle against
    the reference, every data-class output, at 1e-12.  Returns the failures."""
    if HERE not in sys.path:
        sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))       # the repo root, for `oracle`
    from oracle import ctc_oracle as orc
    import tf_numpy_shim
    tf_numpy_shim.install(np.float64)
    if REFERENCE not in sys.path:
        sys.path.append(REFERENCE)
    from tf_seq2seq_losses.classic_ctc_loss import ClassicCtcLossData
    from tf_seq2seq_losses.simplified_ctc_loss import SimplifiedCtcLossData
    rng = np.random.default_rng(seed)
    bad = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for trial in range(trials):
            B, T, V, Lw = int(rng.integers(1, 4)), int(rng.integers(1, 12)), int(rng.integers(2, 7)), int(rng.integers(1, 6))
            blank = int(rng.integers(0, V))
            x = rng.standard_normal((B, T, V)) * rng.choice([0.3, 1.0, 5.0])
            logprobas = x - np.log(np.sum(np.exp(x), axis=2, keepdims=True))
            others = [k for k in range(V) if k != blank]
            labels = rng.choice(others, size=(B, Lw)).astype(np.int32)
            ll = rng.integers(0, Lw + 1, size=B).astype(np.int32)
            tl = rng.integers(0, T + 1, size=B).astype(np.int32)
            for tag, cls, variant in (("classic", ClassicCtcLossData, orc.CLASSIC), ("simplified", SimplifiedCtcLossData, orc.SIMPLIFIED)):
                ref = cls(labels=labels, logprobas=logprobas, label_length=ll, logit_length=tl, blank_index=blank)
                mine = orc.CtcLossData(labels, logprobas, ll, tl, blank, variant)
                for key in ("loss", "gradient", "logarithmic_logproba_gradient", "alpha", "beta", "hessian", "gamma"):
                    got, want = np.asarray(getattr(mine, key), dtype=np.float64), np.asarray(getattr(ref, key), dtype=np.float64)
                    fin = np.isfinite(want)
                    ok = got.shape == want.shape and np.array_equal(got[~fin], want[~fin]) and \
                        (not fin.any() or float(np.max(np.abs(got[fin] - want[fin]))) <= 1e-12)
                    if not ok:
                        bad.append((trial, tag, key, (B, T, V, Lw, blank), ll.tolist(), tl.tolist()))
    return bad


def main():
    if "--fuzz" in sys.argv:
        n = int(sys.argv[sys.argv.index("--fuzz") + 1])
        bad = fuzz(n)
        print(f"{n} random problems x 2 variants x 7 outputs, oracle against the reference at 1e-12; failures: {bad}")
        sys.exit(1 if bad else 0)
    flat = build()
    if "--check" in sys.argv:       # the committed fixture is what the reference computes here: every array, bit for bit
        stored = np.load(OUT)
        bad = [k for k in flat if k not in stored.files or not np.array_equal(np.asarray(flat[k]), stored[k], equal_nan=True)]
        bad += [k for k in stored.files if k not in flat]
        print(f"{len(flat)} arrays regenerated from {REFERENCE}; mismatches: {bad}")
        n, bad2 = check_oracle_made_fixtures()
        print(f"{n} arrays of the oracle-made fixtures tests/golden/*.npz recomputed by the reference; beyond 1e-12: {bad2}")
        sys.exit(1 if bad or bad2 else 0)
    np.savez_compressed(OUT, **flat)
    print(f"wrote {OUT}: {len(flat)} arrays, {os.path.getsize(OUT) / 1024:.0f} KB")


if __name__ == "__main__":
    main()
