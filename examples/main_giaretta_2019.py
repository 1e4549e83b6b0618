"""Giaretta & Girdzijauskas 2019 -- Pegasos gossip on a Barabasi-Albert graph (reference: main_giaretta_2019.py).
``GOSSIPY_NODE_TYPE=passthrough|cacheneigh`` switches to the paper's degree-aware node variants."""
import os

import numpy as np
from _common import cap_nodes, configure, finish, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import PegasosHandler
from gossipy_b200.model.nn import AdaLine
from gossipy_b200.node import CacheNeighNode, GossipNode, PassThroughNode
from gossipy_b200.simul import GossipSimulator, SimulationReport


def barabasi_albert(n: int, m: int, seed: int = 42) -> np.ndarray:
    try:
        from networkx import to_numpy_array
        from networkx.generators.random_graphs import barabasi_albert_graph
        return to_numpy_array(barabasi_albert_graph(n, m, seed=seed)).astype(int)
    except Exception:                                    # preferential attachment without networkx
        rng = np.random.default_rng(seed)
        A = np.zeros((n, n), dtype=int)
        targets = list(range(m))
        repeated = []
        for v in range(m, n):
            for t in set(targets):
                A[v, t] = A[t, v] = 1
            repeated.extend(targets)
            repeated.extend([v] * m)
            targets = list(rng.choice(repeated, size=m))
        return A


rank, world = setup(98765)
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
y = 2 * y - 1
n_train = cap_nodes(int(X.shape[0] * .9))
n_test = X.shape[0] // 10
data_handler = ClassificationDataHandler(X[:n_train + n_test], y[:n_train + n_test], test_size=n_test / (n_train + n_test))
dispatcher = DataDispatcher(data_handler, eval_on_user=False, auto_assign=True)
topology = StaticP2PNetwork(dispatcher.size(), barabasi_albert(dispatcher.size(), min(10, dispatcher.size() - 1)))
model_handler = PegasosHandler(net=AdaLine(data_handler.size(1)), learning_rate=.01,
                               create_model_mode=CreateModelMode.MERGE_UPDATE)
node_cls = {"gossip": GossipNode, "passthrough": PassThroughNode, "cacheneigh": CacheNeighNode}[
    os.environ.get("GOSSIPY_NODE_TYPE", "gossip")]
nodes = node_cls.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                          round_len=100, sync=False)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                      protocol=AntiEntropyProtocol.PUSH, sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(100))
finish(report, rank)
