"""Hegedus et al. 2021 -- partitioned-model gossip with token-account flow control
(reference: main_hegedus_2021.py).  ``GOSSIPY_MODEL=mlp`` runs the BASELINE.json variant
(2-layer MLP on MNIST-shaped data) instead of logistic regression on spambase."""
import os

import torch
from _common import cap_nodes, configure, finish, regular_graph, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
from gossipy_b200.data import DataDispatcher, load_classification_dataset, synthetic
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.flow_control import RandomizedTokenAccount
from gossipy_b200.model.handler import PartitionedTMH
from gossipy_b200.model.nn import LogisticRegression, TorchMLP
from gossipy_b200.model.sampling import TorchModelPartition
from gossipy_b200.node import PartitioningBasedNode
from gossipy_b200.simul import SimulationReport, TokenizedGossipSimulator

rank, world = setup(98765)
n_nodes = cap_nodes(100)
if os.environ.get("GOSSIPY_MODEL", "logreg") == "mlp":
    (Xtr, ytr), (Xte, yte) = synthetic.mnist_like(20000, 2000)
    data_handler = ClassificationDataHandler(Xtr, ytr, Xte, yte)
    net = TorchMLP(784, 10, (100,))
    opt = {"lr": .1, "weight_decay": .001}
else:
    X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
    data_handler = ClassificationDataHandler(X, y, test_size=.1)
    net = LogisticRegression(data_handler.Xtr.shape[1], 2)
    opt = {"lr": 1, "weight_decay": .001}
dispatcher = DataDispatcher(data_handler, n=n_nodes, eval_on_user=False, auto_assign=True)
topology = StaticP2PNetwork(n_nodes, regular_graph(n_nodes, min(20, n_nodes - 1 - (n_nodes - 1) % 2)))
model_handler = PartitionedTMH(net=net, tm_partition=TorchModelPartition(net, 4), optimizer=torch.optim.SGD,
                               optimizer_params=opt, criterion=torch.nn.CrossEntropyLoss(),
                               create_model_mode=CreateModelMode.UPDATE)
nodes = PartitioningBasedNode.generate(data_dispatcher=dispatcher, p2p_net=topology,
                                       model_proto=model_handler, round_len=100, sync=True)
simulator = TokenizedGossipSimulator(nodes=nodes, data_dispatcher=dispatcher,
                                     token_account=RandomizedTokenAccount(C=20, A=10),
                                     utility_fun=lambda mh1, mh2, msg: 1, delta=100,
                                     protocol=AntiEntropyProtocol.PUSH, delay=UniformDelay(0, 10),
                                     sampling_eval=.1)
simulator.native_utility = 1                             # lets the C++ scheduler evaluate the accounts
configure(simulator)
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(1000))
finish(report, rank)
