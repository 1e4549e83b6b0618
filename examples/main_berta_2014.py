"""Berta et al. 2014 -- gossip k-means (reference: main_berta_2014.py)."""
from _common import cap_nodes, configure, finish, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClusteringDataHandler
from gossipy_b200.model.handler import KMeansHandler
from gossipy_b200.node import GossipNode
from gossipy_b200.simul import GossipSimulator, SimulationReport

rank, world = setup(98765)
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
n = cap_nodes(X.shape[0])
data_handler = ClusteringDataHandler(X[:n], y[:n])
dispatcher = DataDispatcher(data_handler, eval_on_user=False, auto_assign=True)
topology = StaticP2PNetwork(dispatcher.size(), None)
model_handler = KMeansHandler(k=2, dim=data_handler.size(1), alpha=.1, matching="hungarian",
                              create_model_mode=CreateModelMode.MERGE_UPDATE)
nodes = GossipNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                            round_len=1000, sync=True)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=1000,
                                      protocol=AntiEntropyProtocol.PUSH, drop_prob=.1, sampling_eval=.01))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(500))
finish(report, rank)
