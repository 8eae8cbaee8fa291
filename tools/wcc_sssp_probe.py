import sys; sys.path.insert(0,'.')
import graph_b200 as gb
g = gb.DiGraph.rmat(24, seed=42, layout=gb.Layout.Sorted)
for _ in range(2): r = g.wcc()
print(g.last_timing())
g2 = gb.DiGraph.rmat(22, seed=42, layout=gb.Layout.Sorted, weights=True)
import numpy as np
off,_ = g2.csr("out"); start = int(np.argmax(np.diff(off.astype(np.int64))))
g2.delta_stepping(start_node=start, delta=0.05); print(g2.last_timing())
