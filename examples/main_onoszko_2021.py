"""Onoszko et al. 2021 -- PENS neighbour selection on rotated CIFAR-10 (reference: main_onoszko_2021.py).
``GOSSIPY_MODEL=resnet20`` swaps the paper's small CNN for ResNet-20 (BASELINE.json config 5)."""
import os

import torch
import torch.nn.functional as F
from _common import cap_nodes, configure, finish, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, CreateModelMode, StaticP2PNetwork
from gossipy_b200.data import DataDispatcher, get_CIFAR10
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import TorchModelHandler
from gossipy_b200.model.nn import CIFAR10Net
from gossipy_b200.node import PENSNode
from gossipy_b200.simul import GossipSimulator, SimulationReport

rank, world = setup(98765)
n_nodes = cap_nodes(8)
n_keep = int(os.environ.get("GOSSIPY_SAMPLES", 4000))
(Xtr, ytr), (Xte, yte) = get_CIFAR10(synthetic_fallback=True)  # (no network: same-shape synthetic data)
Xtr, ytr, Xte, yte = Xtr[:n_keep], torch.as_tensor(ytr[:n_keep]), Xte[:n_keep // 5], torch.as_tensor(yte[:n_keep // 5])
half_tr, half_te = Xtr.shape[0] // 2, Xte.shape[0] // 2
Xtr[half_tr:] = torch.rot90(Xtr[half_tr:], 2, (2, 3))   # second half of the clients sees rotated images
Xte[half_te:] = torch.rot90(Xte[half_te:], 2, (2, 3))
data_handler = ClassificationDataHandler(Xtr, ytr, Xte, yte)
dispatcher = DataDispatcher(data_handler, n=n_nodes, eval_on_user=True, auto_assign=False)
per_tr, per_te = Xtr.shape[0] // n_nodes, Xte.shape[0] // n_nodes
dispatcher.set_assignments([list(range(i * per_tr, (i + 1) * per_tr)) for i in range(n_nodes)],
                           [list(range(i * per_te, (i + 1) * per_te)) for i in range(n_nodes)])
topology = StaticP2PNetwork(n_nodes, None)
if os.environ.get("GOSSIPY_MODEL", "cnn") == "resnet20":
    from gossipy_b200.models import ResNet20
    net = ResNet20(10)
else:
    net = CIFAR10Net()
model_handler = TorchModelHandler(net=net, optimizer=torch.optim.SGD,
                                  optimizer_params={"lr": .01, "weight_decay": .001},
                                  criterion=F.cross_entropy, create_model_mode=CreateModelMode.MERGE_UPDATE,
                                  batch_size=8, local_epochs=int(os.environ.get("GOSSIPY_EPOCHS", 3)))
nodes = PENSNode.generate(data_dispatcher=dispatcher, p2p_net=topology, model_proto=model_handler,
                          round_len=100, sync=False, n_sampled=min(10, n_nodes - 1), m_top=2,
                          step1_rounds=rounds(500) // 5)
simulator = configure(GossipSimulator(nodes=nodes, data_dispatcher=dispatcher, delta=100,
                                      protocol=AntiEntropyProtocol.PUSH, sampling_eval=.1))
report = SimulationReport()
simulator.add_receiver(report)
simulator.init_nodes(seed=42)
simulator.start(n_rounds=rounds(500))
finish(report, rank, local=True)
