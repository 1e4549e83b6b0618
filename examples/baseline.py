"""Centralised baselines (reference: baseline.py): the same model families trained on the pooled data,
as the accuracy ceiling for the gossip curves.  Uses the framework's own fused kernels: one "node"
holding the whole training set."""
import torch
from _common import finish, rounds, setup

from gossipy_b200.core import AntiEntropyProtocol, StaticP2PNetwork
from gossipy_b200.data import DataDispatcher, load_classification_dataset
from gossipy_b200.data.handler import ClassificationDataHandler
from gossipy_b200.model.handler import TorchModelHandler
from gossipy_b200.model.nn import LogisticRegression, TorchMLP

rank, world = setup(98765)
X, y = load_classification_dataset("spambase", as_tensor=True, synthetic_fallback=True)  # (no network: same-shape synthetic data)
data_handler = ClassificationDataHandler(X, y, test_size=.1)
dim = data_handler.Xtr.shape[1]
for name, net, lr in (("logistic regression", LogisticRegression(dim, 2), 1.0),
                      ("MLP (100 hidden)", TorchMLP(dim, 2, (100,)), 0.1)):
    handler = TorchModelHandler(net=net, optimizer=torch.optim.SGD, optimizer_params={"lr": lr, "weight_decay": .001},
                                criterion=torch.nn.CrossEntropyLoss(), local_epochs=1, batch_size=32)
    handler.owner = 0
    handler.init()
    curve = []
    for epoch in range(rounds(10)):
        handler._update(data_handler.get_train_set())
        curve.append(round(handler.evaluate(data_handler.get_eval_set())["accuracy"], 4))
    if rank == 0:
        print("%-22s accuracy per epoch: %s" % (name, curve))
try:
    from sklearn.linear_model import LogisticRegression as SkLR
    Xtr, ytr = data_handler.get_train_set()
    Xte, yte = data_handler.get_eval_set()
    clf = SkLR(max_iter=200).fit(Xtr.numpy(), ytr.numpy())
    if rank == 0:
        print("%-22s accuracy: %.4f" % ("sklearn LogisticRegression", clf.score(Xte.numpy(), yte.numpy())))
except Exception as exc:  # noqa: BLE001
    print("sklearn baseline skipped:", exc)
