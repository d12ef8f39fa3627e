import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import corintho_ai_b200 as cb
from oracle.pyoracle import OracleLib
from diag import compare_trees, _records
O = OracleLib()
fa = cb.fold_batchnorm(cb.random_weights(1))
G, MS, SPE = 16, 32, 8
h = cb.Trainer(G, "", 1, 16, SPE); h.set_weights(fa, 0, "fp32")
for testing in (True, False):
    t = cb.Trainer(G, "", 4, MS, SPE, 1.0, 0.0, 0, 1, testing)
    t.set_weights(fa, 0, "fp32"); 
    if testing: t.set_weights(fa, 1, "fp32")
    t.run_selfplay(2)
    o = O.trainer(num_games=G, seed=4, max_searches=MS, searches_per_eval=SPE, c_puct=1.0, epsilon=0.0, testing=testing)
    ev = np.zeros(G * SPE, np.float32); pr = np.zeros((G * SPE, 96), np.float32); tp = 0 if testing else -1
    o.do_iteration(ev, pr, tp)
    n = o.num_requests(tp); req = o.write_requests(tp); e, p = h.evaluate(req); ev[:n], pr[:n] = e, p
    print("testing", testing, "oracle n", n, "eval[:4]", e[:4], "probs row0 sum", p[0].sum(), p[0][:4])
    o.do_iteration(ev, pr, tp)
    for g in (0, 2):
        print(" game", g, compare_trees(t, O, o, g, 0) or "trees equal")
        eo, ew = t.dump_tree(g, 0)
        recs = _records(ew)
        print("  engine ctl", eo, "root denom bits", hex(int(recs[0][1][5])), "slots[:3]", recs[0][2][:3])
    print(" engine num_requests(tp)", t.num_requests(tp), "oracle", o.num_requests(tp))
