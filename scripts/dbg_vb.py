import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import particulator_b200 as P
import test_gpu_parity as T
co = P.co
tab = T._vb_table()
for sp in (0, 16384):
    ctx = P.Context(device=0)
    ctx.set_option("small_pass_rows", sp)
    e = np.random.default_rng(8).uniform(0, 99.9, 50000) * co.eV
    rg, bg = ctx.table_eval(tab, e)
    print("small_pass", sp, "flags after table_eval", ctx.error_flags(clear=True))
    n = 3000
    rng = np.random.default_rng(4)
    st = dict(x=np.zeros((n, 3)), p=rng.normal(size=(n, 3)) * np.sqrt(2 * 2 * co.eV / co.electron_mass), s=-np.log(1 - rng.random(n)),
              uid=np.arange(1, n + 1, dtype=np.uint64))
    psh = P.RK2Pusher(P.ElectromagneticField(P.HomogeneousField([0.0, 0.0, -100 * co.Td * co.nair]), None))
    ctx.set_rng(6, 0)
    pop = P.Population(ctx, P.SLOW_ELECTRON, 3 * n, st, tab, 0.0)
    mp = P.MultiPopulation(("slow", pop))
    for k in range(3):
        try:
            P.advance(mp, psh, (k + 1) * 1e-12)
        except Exception as ex:
            print("  advance", k, ex)
        d = pop.download()
        E = 0.5 * co.electron_mass * (d["p"] ** 2).sum(1) / co.eV
        print("  step", k, "flags", ctx.error_flags(clear=True), "n", len(E), "Emax", E.max(), "nan", np.isnan(E).sum(), "r range", d["r"].min(), d["r"].max(), P.last_advance_stats(mp)["substeps"])
    ctx.close()
