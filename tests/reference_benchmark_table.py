"""The reference's published benchmark table (tsu/benchmarks/{sampling,optimization,comparison}.py, full
mode) regenerated with the reference's OWN drivers, unmodified.

    python tests/reference_benchmark_table.py                      # drivers bound to the B200 engine (needs a GPU)
    python tests/reference_benchmark_table.py --backend reference  # the same drivers on the reference's own NumPy code

Test-side script (it executes files under oracle/_ref): prints one line per benchmark, the reference's summary() fields."""
import argparse
import contextlib
import importlib.util
import io
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SHIM = os.path.join(ROOT, "tsu_emulator_b200", "compat")


def load_drivers(ref, backend):
    if backend == "b200":
        sys.path.insert(0, SHIM)
        import tsu  # the shim: tsu.gibbs / tsu.core / tsu.models resolve to the engine
    else:
        # the reference package itself, minus its __init__ (which imports matplotlib-based modules)
        tsu = types.ModuleType("tsu")
        tsu.__path__ = [os.path.join(ref, "tsu")]
        sys.modules["tsu"] = tsu
        for name in ("gibbs", "core"):
            spec = importlib.util.spec_from_file_location(f"tsu.{name}", os.path.join(ref, "tsu", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[spec.name] = mod
            spec.loader.exec_module(mod)
    pkg = types.ModuleType("tsu.benchmarks")
    pkg.__path__ = [os.path.join(ref, "tsu", "benchmarks")]
    pkg.__package__ = "tsu.benchmarks"
    sys.modules["tsu.benchmarks"] = pkg
    mods = {}
    for name in ("sampling", "optimization", "comparison"):
        spec = importlib.util.spec_from_file_location(f"tsu.benchmarks.{name}", os.path.join(ref, "tsu", "benchmarks", name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", choices=["b200", "reference"], default="b200")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    from oracle import make_ref

    ref = make_ref.ref_root()
    if ref is None:
        sys.exit("no reference tree (neither /root/reference nor oracle/_ref)")
    mods = load_drivers(ref, args.backend)
    sink = io.StringIO()
    t0 = time.time()
    with contextlib.redirect_stdout(sink):  # the drivers print progress; only the summaries are kept
        samp = mods["sampling"].SamplingBenchmark(seed=42).run_all_benchmarks(quick=args.quick)
        opt = mods["optimization"].OptimizationBenchmark(seed=42).run_all_benchmarks(quick=args.quick)
        cmp_ = mods["comparison"].ComparisonBenchmark(seed=42).run_all_comparisons(quick=args.quick)
    print(f"# reference benchmark drivers, {'quick' if args.quick else 'full'} mode, backend = {args.backend}, "
          f"sampler class = {mods['sampling'].GibbsSampler.__module__}.GibbsSampler, total {time.time() - t0:.1f} s")
    print("# sampling (tsu/benchmarks/sampling.py): name | samples x trials | samples/s | KL | ESS | KS p>0.05")
    for name, r in samp.items():
        s = r.summary()
        print(f"{name} | {s['n_samples']} x {s['n_trials']} | {s['throughput_samples_per_sec']['mean']:.4g} | "
              f"{s['kl_divergence']['mean']:.4g} +- {s['kl_divergence']['std']:.2g} | "
              f"{s['effective_sample_size']['mean']:.4g} | {s['ks_pvalue']['fraction_passed']:.2f}")
    print("# optimisation (tsu/benchmarks/optimization.py): problem | size | trials | best objective mean (best) | ms per problem")
    for name, r in opt.items():
        s = r.summary()
        print(f"{name} | {s['size']} | {s['n_trials']} | {s['best_objective']['mean']:.4g} ({s['best_objective']['best']:.4g}) | "
              f"{s['solution_time_ms']['mean']:.4g}")
    print("# comparison (tsu/benchmarks/comparison.py): problem | framework | objective mean | ms")
    for name, r in cmp_.items():
        s = r.summary()
        for fw, v in s["frameworks"].items():
            print(f"{name} | {fw} | {v['objective']['mean']:.4g} | {v['time_ms']['mean']:.4g}")


if __name__ == "__main__":
    main()
