"""Ormandi et al. 2013 -- gossip learning with Pegasos linear models (reference: main_ormandi_2013.py).
One training sample per node, clique, PUSH, delays / churn / message loss."""
from _common import cap_nodes, configure, finish, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork, UniformDelay
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import PegasosHandler
from gossipy_b200.model.nn import AdaLine
from gossipy_b200.node import GossipNode
from gossipy_b200.simul import GossipSimulator, SimulationReport

rank, world = setup(98765)
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
y = 2 * y - 1                                            # labels in {-1, +1}
n_train = cap_nodes(int(X.shape[0] * .9))
data_handler = ClassificationDataHandler(X[:n_train + X.shape[0] // 10], y[:n_train + X.shape[0] // 10],
                                         test_size=X.shape[0] // 10 / (n_train + X.shape[0] // 10))
dispatcher = DataDispatcher(data_handler, eval_on_user=False, auto_assign=True)   # one sample per node
topology = StaticP2PNetwork(dispatcher.size(), None)
model_handler = PegasosHandler(net=AdaLine(data_handler.size(1)), learning_rate=.01,
                               create_model_mode=CreateModelMode.MERGE_UPDATE)
nodes = GossipNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                            round_len=100, sync=False)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                      protocol=AntiEntropyProtocol.PUSH, delay=UniformDelay(0, 10),
                                      online_prob=.2, drop_prob=.1, sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(100))
finish(report, rank)
