"""Koloskova et al. 2020 -- decentralised SGD with neighbourhood averaging (reference: main_all2all.py).
``GOSSIPY_SYNC=1`` runs synchronous rounds on a clique: one NVLS / P2P all-reduce kernel per round."""
import os

import torch
from _common import cap_nodes, configure, finish, regular_graph, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformMixing
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import WeightedTMH
from gossipy_b200.model.nn import LogisticRegression
from gossipy_b200.node import All2AllGossipNode
from gossipy_b200.simul import All2AllGossipSimulator, SimulationReport

rank, world = setup(98765)
sync_rounds = os.environ.get("GOSSIPY_SYNC", "0") == "1"
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
data_handler = ClassificationDataHandler(X, y, test_size=.1)
n_nodes = cap_nodes(8 if sync_rounds else 100)
dispatcher = DataDispatcher(data_handler, n=n_nodes, eval_on_user=False, auto_assign=True)
adj = None if sync_rounds else regular_graph(n_nodes, min(20, n_nodes - 1 - (n_nodes - 1) % 2))
topology = StaticP2PNetwork(n_nodes, adj)
net = LogisticRegression(data_handler.Xtr.shape[1], 2)
model_handler = WeightedTMH(net=net, optimizer=torch.optim.SGD, optimizer_params={"lr": .1, "weight_decay": .01},
                            criterion=torch.nn.CrossEntropyLoss(), create_model_mode=CreateModelMode.MERGE_UPDATE)
nodes = All2AllGossipNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                                   round_len=100, sync=False)
simulator = configure(All2AllGossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                             protocol=AntiEntropyProtocol.PUSH, sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(UniformMixing(topology), n_rounds=rounds(100), synchronous=sync_rounds)
finish(report, rank)
