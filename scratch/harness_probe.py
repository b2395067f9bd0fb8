"""Run the unchanged simple_fusion.py in a scratch directory a few times and print the cross-validated C-indices (seeded / unseeded)."""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "harness"))
import ref_cohort
for seed in sys.argv[1:]:
    d = tempfile.mkdtemp()
    ref_cohort.write(d)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "shim"), ROOT, os.path.join(ROOT, "tests", "harness", "stubs")])
    env["PYTHONIOENCODING"] = "utf-8"
    if seed != "none":
        env["B200SURV_HARNESS_SEED"] = seed
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_ref_scripts", "simple_fusion.py")], cwd=d, env=env, capture_output=True, text=True)
    try:
        cv = json.load(open(os.path.join(d, "results", "simple_fusion", "cv_results.json")))
        print(seed, "rc", r.returncode, "mean", round(cv["c_index_mean"], 4), [round(f["best_c_index"], 4) for f in cv["fold_results"]], flush=True)
    except Exception as e:
        print(seed, "rc", r.returncode, "no cv:", e, r.stderr[-500:], flush=True)
