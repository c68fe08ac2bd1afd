"""Golden vectors of calculate_optimum (src/rl_opt.py:26-152) from the UNMODIFIED reference: the full 24-column
statistics of two small splits.  Run where the reference is mounted:  python tests/golden/gen_golden_topt.py"""
import contextlib, io, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_harness import ReferenceSession  # noqa: E402

out = {}
meta = []
for name, overrides, split in (("bs2_op2_test", dict(scenario=2, operation="OP2"), "test"),
                               ("bs3_op1_val", dict(scenario=3, operation="OP1"), "val"),
                               ("bs1_op1_test", dict(scenario=1, operation="OP1"), "test")):
    sess = ReferenceSession(overrides)
    import src.rl_opt as ro
    old = os.getcwd(); os.chdir(sess.tmp)
    with contextlib.redirect_stdout(io.StringIO()):
        d = ro.calculate_optimum(sess.dict_price_data[f"el_price_{split}"], sess.dict_price_data[f"gas_price_{split}"],
                                 sess.dict_price_data[f"eua_price_{split}"], "Test_set", sess.E.stats_names)
    os.chdir(old)
    out[name] = np.stack([np.asarray(d[k], dtype=np.float64) for k in sess.E.stats_names], axis=1)
    meta.append(dict(name=name, overrides=overrides, split=split, stats_names=list(sess.E.stats_names)))
    sess.close()
np.savez_compressed(os.path.join(HERE, "topt_reference.npz"), meta=json.dumps(meta), **out)
print({k: v.shape for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "topt_reference.npz")) // 1024, "KiB")
