import sys; sys.path.insert(0, ".")
import numpy as np
from mogpe_b200._model_core import Adam
from tests._model_helpers import model_from_spec
from tests._specs import make_spec
spec, X, Y, eps = make_spec(seed=85, N=100, D=2, F=1, K=2, M=8, S=2, num_data=100)
def run(graphs, split):
    m = model_from_spec(spec); m.set_seed(11); m.compile(optimizer=Adam(learning_rate=0.01)); m._core.use_graphs = graphs
    if split:
        h = m.fit(X, Y, epochs=2, batch_size=32, shuffle=False)["loss"]; m.predict_y(X[:10]); h += m.fit(X, Y, epochs=2, batch_size=32, shuffle=False)["loss"]
    else:
        h = m.fit(X, Y, epochs=4, batch_size=32, shuffle=False)["loss"]
    print("graphs", graphs, "split", split, h, m._core.engine.graph_stats())
    return h
run(False, False); run(False, True); run(True, False); run(True, True)
