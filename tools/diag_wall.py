"""Where the write phase of the drop-in driver spends its time (C4): timers around the writer and the engine's close."""
import contextlib, io, os, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wgdi_b200 import base, synth, run_colliearity, engine, _host

w = sys.argv[1] if len(sys.argv) > 1 else "C4"
d = tempfile.mkdtemp(prefix="diag_", dir="/dev/shm")
T = {}


def timed(name, fn):
    def f(*a, **k):
        t = time.perf_counter()
        try:
            return fn(*a, **k)
        finally:
            T[name] = T.get(name, 0.0) + time.perf_counter() - t
    return f


run_colliearity.write_blocks_native = timed("write_blocks_native", run_colliearity.write_blocks_native)
run_colliearity._native_gene_tables = timed("gene_tables", run_colliearity._native_gene_tables)
_host.write_collinearity = timed("write_collinearity", _host.write_collinearity)
engine.IclEngine.close = timed("engine.close", engine.IclEngine.close)
engine.IclEngine.chain = timed("engine.chain", engine.IclEngine.chain)
engine.IclEngine.__init__ = timed("engine.init", engine.IclEngine.__init__)
run_colliearity.compact_points = timed("compact_points", run_colliearity.compact_points)
try:
    ds = synth.make_dataset(w)
    conf = synth.write_dataset(ds, d)
    for mode in ("native", "native", "native-steps", "native"):
        T.clear()
        opts = list(base.load_conf(conf, "collinearity")) + [("ingest", mode)]
        with contextlib.redirect_stdout(io.StringIO()):
            m = run_colliearity.mycollinearity(opts)
            t0 = time.perf_counter()
            try:
                m.run()
            except SystemExit:
                pass
            dt = time.perf_counter() - t0
        print(mode, "wall %.3f" % dt, {k: round(v, 3) for k, v in m.timing.items() if k.endswith("_s")}, {k: round(v, 3) for k, v in T.items()}, flush=True)
finally:
    shutil.rmtree(d, ignore_errors=True)
