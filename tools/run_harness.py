#!/usr/bin/env python
"""Drive the B200 plugin through the reference's own harness, offline.

    python tools/run_harness.py --stage                       # copy /root/reference -> baseline/_ref (git-ignored)
    python tools/run_harness.py --prepare NAME [NAME ...]     # write dataset files (ground truth = reference brute force)
    python tools/run_harness.py --dataset NAME --hardware GPU # run.py --local --definitions <repo>/algos.yaml ...
    python tools/run_harness.py --dataset NAME --hardware CPU --algorithm bruteforce-product-blas
    python tools/run_harness.py --score NAME [--json out.jsonl]
    python tools/run_harness.py --list

``--dataset`` calls the reference's ``main()`` unmodified (main.py:74-308): it parses algos.yaml,
forks its worker, ``runner.run`` times ``fit()`` / ``query()`` with its host clock and
``results.store_result`` writes ``results/<dataset>/<algo>/<args>.hdf5`` under the reference tree.
``--score`` reads those files back with ``results.load_all_results`` and evaluates every entry of the
reference's ``plotting.metrics.all_metrics`` (``plotting.utils.compute_all_metrics``), plus the relative
L2 error BASELINE.json's tolerance is stated in (the reference only has absolute errors) and, for solver
datasets, the relative residual of the system.
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from kernel_matrix_benchmarks_b200.harness import bootstrap  # noqa: E402


def score(dataset, json_path=None):
    root, _ = bootstrap.activate()
    import numpy as np
    from kernel_matrix_benchmarks.datasets import get_dataset
    from kernel_matrix_benchmarks.plotting.utils import compute_all_metrics
    from kernel_matrix_benchmarks.results import load_all_results

    cwd = os.getcwd()
    os.chdir(root)
    rows = []
    try:
        ds, _ = get_dataset(dataset)
        for properties, run in load_all_results(dataset):
            scored = compute_all_metrics(dataset=ds, run=run, properties=properties, recompute=True)
            result, error = run["result"][:], run["error"][:]
            truth = result - error
            row = {"dataset": dataset, "algo": scored["algo"], "name": scored["algo_name"]}
            row.update({k: float(v) for k, v in scored["metrics"].items()})
            row["rel-l2-error"] = float(np.linalg.norm(error) / max(np.linalg.norm(truth), 1e-300))
            if ds.attrs["task"] == "solver":
                # the residual metric plotting/utils.py:83-86 anticipates: |(K + lam I) x - a| / |a| with the
                # reference's float64 brute force as K (the stored error compares with the generating b, which is
                # meaningless when K is singular to working precision -- SURVEY.md section 8c)
                from kernel_matrix_benchmarks_b200.harness.datasets_ext import VERIFY_ROWS, _ground_truth_blocked

                lam = float(ds.attrs.get("lam", 0.0))
                a = ds["target_signal"][:]
                pts = ds["source_points"][:]
                n = pts.shape[0]
                # beyond ~10^9 pairs the residual is taken on sampled rows (each row still sees every source)
                srows = np.arange(n) if float(n) * n <= 2e9 else np.sort(np.random.RandomState(13).choice(n, VERIFY_ROWS, replace=False))
                Kx = _ground_truth_blocked(kernel=ds.attrs["kernel"], source_points=pts, target_points=np.ascontiguousarray(pts[srows]),
                                           source_signal=result, normalize_rows=False)
                row["rel-residual"] = float(np.linalg.norm(Kx + lam * result[srows] - a[srows]) / max(np.linalg.norm(a[srows]), 1e-300))
                row["rel-residual-rows"] = int(len(srows))
            for k, v in properties.items():
                if k not in row and isinstance(v, (int, float, str, bool)):
                    row[k] = v
            rows.append(row)
        ds.close()
    finally:
        os.chdir(cwd)
    rows.sort(key=lambda r: r["total-time"])
    for r in rows:
        print(json.dumps(r))
    if json_path:
        with open(json_path, "a") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")
    return rows


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--prepare", nargs="+", metavar="NAME")
    ap.add_argument("--dataset")
    ap.add_argument("--hardware", default="GPU", choices=["CPU", "GPU"])
    ap.add_argument("--algorithm", default=None)
    ap.add_argument("--definitions", default=None, help="default: this repo's algos.yaml (GPU) / the reference's (CPU)")
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--score", metavar="NAME")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()

    if args.stage:
        dst = bootstrap.stage_reference()
        print("staged reference at", dst)
    if args.list:
        bootstrap.activate()
        from kernel_matrix_benchmarks.datasets import DATASETS

        print("\n".join(DATASETS.keys()))
    if args.prepare:
        root, shimmed = bootstrap.activate()
        from kernel_matrix_benchmarks.datasets import DATASETS, get_dataset_fn

        cwd = os.getcwd()
        os.chdir(root)
        try:
            for name in args.prepare:
                fn = get_dataset_fn(name)  # data/<name>.hdf5 (datasets.py:94-98)
                if os.path.exists(fn):
                    print("exists:", os.path.join(root, fn))
                else:
                    DATASETS[name](fn)
        finally:
            os.chdir(cwd)
    if args.dataset:
        root = bootstrap.find_reference()
        definitions = args.definitions or (
            os.path.join(REPO, "algos.yaml") if args.hardware == "GPU" else os.path.join(root, "algos.yaml"))
        argv = ["--local", "--force", "--hardware", args.hardware, "--definitions", definitions,
                "--dataset", args.dataset, "--runs", str(args.runs)]
        if args.algorithm:
            argv += ["--algorithm", args.algorithm]
        bootstrap.run_main(argv)
    if args.score:
        score(args.score, args.json)


if __name__ == "__main__":
    main()
