"""Generate tests/golden/evaluate.npz by running the UNMODIFIED reference harness: predict.evaluate (predict.py:103-135)
and predict.benchmark (predict.py:138-158) on the tiny datasets of oracle/cases.py (EVALUATE_SETS).

Build container only (needs /root/reference):    python oracle/make_golden_evaluate.py

The reference's evaluate() only runs inside IPython: it reads the global ``__IPYTHON__`` and rebinds ``tqdm`` as a local
(predict.py:127-128), so outside a notebook the loop header raises.  The script therefore defines ``__IPYTHON__ = True`` in
builtins and replaces ``tqdm_notebook`` by a pass-through; the imagesets are DataLoader.ImageSet objects like the reference's own.
HR / mask of an imageset: built from the reference SR of that imageset (roll + 0.02 + sigma = 0.01 noise, mask 90 % clear)
so that the score is well conditioned; both are stored because tests cannot regenerate them without the reference."""
from __future__ import annotations

import builtins
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import cases, hrnet_oracle, predict_oracle  # noqa: E402
from oracle.make_golden import _import_reference  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    HRNet, _, _ = _import_reference()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import predict as ref_predict  # type: ignore
        from DataLoader import ImageSet  # type: ignore
    builtins.__IPYTHON__ = True
    ref_predict.tqdm_notebook = lambda it, *a, **k: it
    with open("/root/reference/config/config.json") as f:
        ref_cfg = json.load(f)["network"]
    params = hrnet_oracle.make_params(cases.WEIGHT_SEED)
    model = HRNet(ref_cfg).eval()
    model.load_state_dict(params, strict=True)

    out = {}
    rng = np.random.RandomState(78)
    datasets = {"train": [], "val": [], "test": []}
    for name, (case, split) in cases.EVALUATE_SETS.items():
        n, s, has_hr = cases.PREDICT_CASES[case]
        lr = cases.evaluate_lrs(name)
        imset = {"name": name, "lr": torch.from_numpy(lr), "hr": None, "hr_map": torch.ones(3 * s, 3 * s),
                 "clearances": cases.evaluate_clearances(name)}
        if has_hr:
            with torch.no_grad():
                sr0, _ = ref_predict.get_sr_and_score(ImageSet(imset), model, min_L=cases.PREDICT_MIN_L)
            hr = np.clip(np.roll(np.clip(sr0, 0, 1), (-1, 2), (0, 1)) + 0.02 + 0.01 * rng.randn(*sr0.shape), 0, 1).astype(np.float32)
            hm = (rng.rand(*sr0.shape) > 0.1).astype(np.float32)
            imset["hr"], imset["hr_map"] = torch.from_numpy(hr), torch.from_numpy(hm)
            out[name + "__hr"], out[name + "__hr_map"] = hr, hm
        datasets[split].append(ImageSet(imset))
    with torch.no_grad(), np.errstate(all="ignore"):
        scores, clerances, part = ref_predict.evaluate(model, datasets["train"], datasets["val"], datasets["test"],
                                                       min_L=cases.PREDICT_MIN_L)
    assert list(scores) == list(cases.EVALUATE_SETS)
    for name in scores:
        split = cases.EVALUATE_SETS[name][1]
        assert part[name] == split and np.array_equal(clerances[name], cases.evaluate_clearances(name))
        out[name + "__score"] = np.float32(np.nan if scores[name] is None else scores[name])
        # the oracle restatement agrees with the reference harness
        im = {"lr": cases.evaluate_lrs(name), "hr": out.get(name + "__hr"), "hr_map": out.get(name + "__hr_map")}
        _, o_score = predict_oracle.get_sr_and_score(im, params, cases.PREDICT_MIN_L)
        assert (scores[name] is None and o_score is None) or abs(float(scores[name]) - float(o_score)) <= 1e-4
        print(f"evaluate {name} [{split}]: score {scores[name]} oracle {o_score}")
    # benchmark on the scored splits (the reference divides by the model score, so the unscored test set is left out)
    scored = {k: v for k, v in scores.items() if v is not None}
    table = ref_predict.benchmark({k: cases.EVALUATE_BASELINE[k] for k in scored}, scored,
                                  {k: part[k] for k in scored}, {k: clerances[k] for k in scored})
    out["benchmark__index"] = np.array(list(table.index))
    for col in ("ESA", "model", "score", "mean_clr", "std_clr"):
        out["benchmark__" + col] = table[col].to_numpy(dtype=np.float64)
    out["benchmark__part"] = np.array(list(table["part"]))
    out["benchmark__columns"] = np.array(list(table.columns))
    print(table)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "evaluate.npz"), **out)


if __name__ == "__main__":
    main()
