"""The host time loop — mirrors src/run.jl:1-40."""
import logging

from .mixed_population import advance
from .population import droplow, nparticles, spread
from .callback import VoidCallback

log = logging.getLogger("particulator_b200")


def _msg(mpopl, t, perc):
    """run.jl:31-40"""
    nstr = "\n".join(f"{str(k):>15s} => {nparticles(p)!r:<10}" for k, p in mpopl.pairs())
    lstr = "\n".join(f"{str(k):>15s} => {spread(p)[0]!r:<10}" for k, p in mpopl.pairs())
    return f"{perc:.2f}% complete\ntime = {t / 1e-9:.3f} ns\n# of particles:\n{nstr}\ncentroid locations:\n{lstr}"


def run(mpopl, pusher, tfinal, dt, callback=None, output_dt="default", verbosity=1):
    """run!(mpopl, pusher, tfinal, dt, callback; output_dt=tfinal/20, verbosity=1)  run.jl:1-29"""
    callback = callback or VoidCallback()
    if output_dt == "default":
        output_dt = tfinal / 20
    t = 0.0
    nxt = t
    isave = 0
    while t < tfinal:
        advance(mpopl, pusher, t + dt, callback)
        for popl in mpopl:
            droplow(popl)
        t += dt
        cont = callback.onstep(mpopl, t)
        if output_dt is not None and t >= nxt:
            nxt += output_dt
            isave += 1
            if verbosity > 0:
                log.info(_msg(mpopl, t, 100 * t / tfinal))
            callback.onoutput(mpopl, t, isave)
        if not cont:
            if verbosity > 0:
                log.info("Early stop due to onstep() callback")
            break
    return t
