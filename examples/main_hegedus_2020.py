"""Hegedus et al. 2020 -- decentralised matrix factorisation (reference: main_hegedus_2020.py)."""
from _common import cap_nodes, configure, finish, regular_graph, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
from gossipy_b200.data import RecSysDataDispatcher, load_recsys_dataset
from gossipy_b200.data.handler import RecSysDataHandler
from gossipy_b200.model.handler import MFModelHandler
from gossipy_b200.node import GossipNode
from gossipy_b200.simul import GossipSimulator, SimulationReport

rank, world = setup(98765)
ratings, n_users, n_items = load_recsys_dataset("ml-1m", synthetic_fallback=True)  # (no network: same-shape synthetic data)
n_users = cap_nodes(n_users)
ratings = {u: r for u, r in ratings.items() if u < n_users}
data_handler = RecSysDataHandler(ratings, n_users, n_items, test_size=.1, seed=42)
dispatcher = RecSysDataDispatcher(data_handler)
topology = StaticP2PNetwork(n_users, regular_graph(n_users, min(20, n_users - 1 - (n_users - 1) % 2)))
model_handler = MFModelHandler(dim=5, n_items=n_items, lam_reg=.1, learning_rate=.001,
                               create_model_mode=CreateModelMode.MERGE_UPDATE)
nodes = GossipNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                            round_len=100, sync=True)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                      protocol=AntiEntropyProtocol.PUSH, delay=UniformDelay(0, 10),
                                      sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(100))
finish(report, rank, local=True)
