"""`main.py`-equivalent experiment loop with the stacked client gradients resident on the GPU.

Reference behaviour mirrored (file:line in /root/reference):
    main(...)                      main.py:12-100   build users + server, epoch loop, test every 5 epochs, CSV, checkpoint
    corrupted_count, is_malicious  main.py:21,28    malicious ids are 0..int(mal_prop*N)-1
    epoch body                     main.py:64-71    dispatch_weights -> attacker.attack(mal_users) -> collect -> defend
    learning-rate fading           server.py:50-52  lr_t = lr * fading / (epoch + fading) for the CLIENT optimiser only;
                                                    the server step uses the base rate (server.py:89, reproduced)
    test                           server.py:92-112 sum of per-batch NLL / dataset size, argmax accuracy
    outputs                        main.py:85-89    torch.save({'epoch','state_dict','acc'}) to runs/<dataset>/checkpoint.pth.tar
                                   main.py:100      np.savetxt('logs/<...>.csv', accuracies, delimiter=',')

What differs, on purpose: (i) the datasets need a download (data_sets.py:30) that is impossible offline, so the clients
train on a seeded synthetic 10-class 28x28 problem with the reference's MnistNet architecture (data_sets.py:13-24,
D = 79,510); (ii) clients are evaluated on the GPU and write their flat gradients straight into their row of the
device-resident N x D matrix (ingest.ParamLayout: user.py:17-28 order), so attack, defence and the server step never
leave the device.  The training simulation itself is not part of the accelerated path.

    python -m attacking_federate_learning_b200.harness -d Krum -e 30 --users-count 10
"""
from __future__ import annotations

import argparse
import datetime
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import malicious
from .ingest import ParamLayout
from .server import AggregationServer

SYNTH = 'SYNTH-MNIST'


class MnistNet(nn.Module):                       # data_sets.py:13-24
    def __init__(self):
        super().__init__()
        self.fc1 = nn.Linear(28 * 28, 100)
        torch.nn.init.xavier_uniform_(self.fc1.weight)
        self.fc2 = nn.Linear(100, 10)

    def forward(self, x):
        return F.log_softmax(self.fc2(F.relu(self.fc1(x))), dim=1)


def synthetic_problem(n_train, n_test, device, seed=0):
    """10 Gaussian class prototypes in 784 dimensions + noise: learnable to > 90 % by the MLP in a few dozen rounds."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    protos = torch.randn(10, 28 * 28, generator=g)

    def draw(n):
        y = torch.randint(0, 10, (n,), generator=g)
        x = protos[y] + 2.5 * torch.randn(n, 28 * 28, generator=g)
        return x.to(device), y.to(device)
    return draw(n_train), draw(n_test)


class Client:
    """user.py:33-92 without the data loader machinery: one minibatch forward/backward per round, flat gradient out."""

    def __init__(self, user_id, is_malicious, x, y, batch_size, layout, device):
        self.user_id, self.is_malicious = user_id, is_malicious
        self.x, self.y, self.batch_size, self.pos = x, y, batch_size, 0
        self.net = MnistNet().to(device)
        self.layout = layout
        self.criterion = nn.NLLLoss()
        self.grads = None
        self.original_params = None
        self.learning_rate = None

    def step(self, current_params, learning_rate, out_row):
        if self.user_id == 0 and self.is_malicious:                       # user.py:84-86
            self.original_params = current_params.clone()
            self.learning_rate = learning_rate
        params = list(self.net.parameters())
        self.layout.row_into_parameters(current_params, params)           # user.py:87
        lo = self.pos
        hi = min(lo + self.batch_size, len(self.x))
        self.pos = 0 if hi == len(self.x) else hi                         # cycle(train_loader)
        self.net.zero_grad(set_to_none=True)
        loss = self.criterion(self.net(self.x[lo:hi]), self.y[lo:hi])
        loss.backward()                                                   # no optimiser step: the server steps (user.py:81)
        self.layout.flatten([p.grad for p in params], out=out_row)        # user.py:92, written into the matrix row
        self.grads = out_row


def main(mal_prop, num_std, defense, users_count=10, epochs=150, learning_rate=0.1, fading_rate=10000, momentum=0.9,
         batch_size=83, output=None, device="cuda", out_dir=".", seed=0, train_size=20000, test_size=4000, test_step=5):
    if output:
        def my_print(s, end='\n'):
            with open(output, 'a+') as f:
                f.write(str(s) + end)
    else:
        my_print = print
    my_print(dict(mal_prop=mal_prop, num_std=num_std, defense=defense, users_count=users_count, epochs=epochs,
                  learning_rate=learning_rate, dataset=SYNTH))
    torch.manual_seed(seed)
    corrupted_count = int(mal_prop * users_count)                         # main.py:21
    (xtr, ytr), (xte, yte) = synthetic_problem(train_size, test_size, device, seed)
    test_net = MnistNet().to(device)
    layout = ParamLayout(test_net.parameters())
    srv = AggregationServer(users_count, layout.dim, mal_prop, learning_rate, momentum, device=device,
                            initial_weights=layout.flatten(list(test_net.parameters())))
    users = [Client(u, u < corrupted_count, xtr[u::users_count], ytr[u::users_count], batch_size, layout, device)
             for u in range(users_count)]                                 # DistributedSampler-style partition (user.py:50)
    attacker = malicious.DriftAttack(num_std)
    my_print("\nStarting Training...")
    criterion = nn.NLLLoss()
    accuracies, accuracies_epochs = [], []
    for epoch in range(epochs):
        lr_t = learning_rate * fading_rate / (epoch + fading_rate)       # server.py:50-52
        for u in users:                                                   # server.py:54-56 dispatch_weights
            u.step(srv.current_weights, lr_t, srv.users_grads[u.user_id])
        attacker.attack_rows(srv.users_grads, corrupted_count)            # main.py:66-68 + server.py:82-83, in place
        srv.defend(defense, epoch)                                        # main.py:71
        if epoch % test_step == 0 or epoch == epochs - 1:
            layout.row_into_parameters(srv.current_weights, list(test_net.parameters()))
            test_net.eval()
            test_loss, correct = 0.0, 0
            with torch.no_grad():
                for lo in range(0, len(xte), batch_size):                 # server.py:100-110
                    out = test_net(xte[lo:lo + batch_size])
                    test_loss += criterion(out, yte[lo:lo + batch_size]).item()
                    correct += int(out.max(1)[1].eq(yte[lo:lo + batch_size]).sum())
            test_loss /= len(xte)
            accuracy = 100. * float(correct) / len(xte)
            my_print('Test set: [{:3d}] Average loss: {:.4f}, Accuracy: {}/{} ({:.2f}%)'.format(epoch, test_loss, correct,
                                                                                             len(xte), accuracy))
            accuracies.append(accuracy)
            accuracies_epochs.append(epoch)
            if accuracy > 70.:                                            # main.py:84-89, server.py:40-46
                directory = os.path.join(out_dir, "runs", SYNTH)
                os.makedirs(directory, exist_ok=True)
                torch.save({'epoch': epoch + 1, 'state_dict': test_net.state_dict(), 'acc': accuracy},
                           os.path.join(directory, 'checkpoint.pth.tar'))
    my_print(datetime.datetime.now().time())
    my_print("Max accuracy: {}".format(max(accuracies)))
    os.makedirs(os.path.join(out_dir, "logs"), exist_ok=True)             # the reference needs a pre-made logs/ (readme.md:25)
    csv = os.path.join(out_dir, 'logs', '{}_stdev_{}_{}_backdoor-{}_mal_prop_{}_users_{}_alpha_{}_lr_{}.csv'.format(
        SYNTH, num_std, defense, False, mal_prop, users_count, None, learning_rate))
    np.savetxt(csv, accuracies, delimiter=',')                            # main.py:100
    return accuracies, accuracies_epochs, csv


if __name__ == '__main__':
    p = argparse.ArgumentParser()                                         # main.py:104-131 (the flags that apply here)
    p.add_argument('-m', '--mal-prop', default=0.24, type=float)
    p.add_argument('-z', '--num_std', default=1.5, type=float)
    p.add_argument('-d', '--defense', default='NoDefense', choices=['NoDefense', 'Bulyan', 'TrimmedMean', 'Krum'])
    p.add_argument('-n', '--users-count', default=10, type=int)
    p.add_argument('-c', '--batch_size', default=128, type=int)
    p.add_argument('-e', '--epochs', default=300, type=int)
    p.add_argument('-l', '--learning_rate', default=0.1, type=float)
    p.add_argument('-o', '--output', type=str)
    a = p.parse_args()
    main(a.mal_prop, a.num_std, a.defense, users_count=a.users_count, epochs=a.epochs, learning_rate=a.learning_rate,
         batch_size=a.batch_size, output=a.output)
