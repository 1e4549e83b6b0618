"""Danner et al. 2023 -- age-limited merge under churn (reference: main_danner_2023.py)."""
import torch
from _common import cap_nodes, configure, finish, regular_graph, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import LimitedMergeTMH
from gossipy_b200.model.nn import LogisticRegression
from gossipy_b200.node import GossipNode
from gossipy_b200.simul import GossipSimulator, SimulationReport

rank, world = setup(98765)
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
data_handler = ClassificationDataHandler(X, y, test_size=.1)
n_nodes = cap_nodes(100)
dispatcher = DataDispatcher(data_handler, n=n_nodes, eval_on_user=False, auto_assign=True)
topology = StaticP2PNetwork(n_nodes, regular_graph(n_nodes, min(20, n_nodes - 1 - (n_nodes - 1) % 2)))
net = LogisticRegression(data_handler.Xtr.shape[1], 2)
model_handler = LimitedMergeTMH(net=net, optimizer=torch.optim.SGD,
                                optimizer_params={"lr": 1, "weight_decay": .001},
                                criterion=torch.nn.CrossEntropyLoss(),
                                create_model_mode=CreateModelMode.MERGE_UPDATE, age_diff_threshold=1)
nodes = GossipNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                            round_len=100, sync=True)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                      protocol=AntiEntropyProtocol.PUSH, delay=UniformDelay(0, 10),
                                      online_prob=.2, drop_prob=.1, sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(1000))
finish(report, rank)
