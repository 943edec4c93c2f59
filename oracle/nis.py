"""Oracle: host formulas of BasicManager (variance loss, integrate combine, epoch state machine).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates
  /root/reference/nisrep/normalizing_flows/manager.py:66-378  (_train_variance_forward_seq)
  /root/reference/nisrep/normalizing_flows/manager.py:380-405 (integrate)
"""
import math

import torch


def minibatch_loss(fres, J, maxf, loss_mode="var"):
    """manager.py:237-258 — loss contribution of one minibatch.  fres = f(w) (preburn) or f(X)."""
    fXJ = fres * J / maxf
    if loss_mode == "var":
        return torch.var(fXJ)                    # unbiased
    if loss_mode == "est":
        return torch.mean((fXJ * maxf) ** 2)
    raise ValueError("Unknown loss function")


def integrate_combine(mean, var, neval):
    """manager.py:402-403.  mean, var: [nitn].  NOTE the reference's error divides by
    sqrt(neval*nitn) after harmonic-summing nitn variances, i.e. it is ~sqrt(nitn) too small;
    ``honest_err`` is the correct standard error of the inverse-variance-weighted mean."""
    nitn = mean.shape[0]
    sig = torch.sum(mean / var) / torch.sum(1 / var)
    sig_err = torch.sqrt(1 / torch.sum(1 / var)) / math.sqrt(neval * nitn)
    honest_err = torch.sqrt(1 / torch.sum(neval / var))
    return sig, sig_err, honest_err


def integrate(model_fn, f, nitn, neval, n_flow, generator):
    """manager.py:380-405 with an explicit generator.  model_fn maps [B,d+1] -> [B,d+1]."""
    mean = torch.zeros(nitn, dtype=torch.float64)
    var = torch.zeros(nitn, dtype=torch.float64)
    for i in range(nitn):
        w = torch.rand(neval, n_flow, generator=generator, dtype=torch.float32).double()   # :390,:395
        Y = torch.cat((w, torch.ones(neval, 1, dtype=torch.float64)), 1)                   # AddJacobian
        X = model_fn(Y).detach()
        fres = f(X[:, :-1]) * X[:, -1]
        var[i] = torch.var(fres)
        mean[i] = torch.mean(fres)
    return integrate_combine(mean, var, neval)


class EpochStateMachine:
    """manager.py:205-327 — best-model bookkeeping, preburn switch, early stop, as a pure function
    of the per-epoch loss sequence.  ``step(i, loss)`` returns True when training must stop."""

    def __init__(self, int_loss, preburn_time=75, kill_counter=7, impr_ratio=1e-2, save_best=True):
        self.check_time = preburn_time if preburn_time > 10 else 50       # :78-81
        self.preburn_time = preburn_time
        self.kill_counter = kill_counter
        self.impr_ratio = impr_ratio
        self.save_best = save_best
        self.int_loss = int_loss
        self.best_loss = int_loss
        self.best_epoch = 0
        self.stale_save = 1000
        self.preburner = preburn_time > 0
        self.counter = 0
        self.last_loss = 1000
        self.snapshots = []          # epochs at which best_model was deep-copied

    def step(self, i, loss):
        stop = False
        if self.save_best and loss < self.best_loss and not self.preburner:       # :293-298
            self.best_loss = loss
            self.best_epoch = i
            self.snapshots.append(i)
        if loss < self.last_loss:                                                  # :307-315
            self.counter = 0
        else:
            self.counter += 1
            if self.counter > self.kill_counter and self.preburner:
                self.counter = 0
                self.preburner = False
            elif self.counter > self.kill_counter:
                stop = True
        if not stop:
            self.last_loss = loss
            if (i % self.check_time == 0) and i > (self.preburn_time + 1) and \
                    float(self.best_loss / self.stale_save) > (1 - self.impr_ratio) and not self.preburner:
                stop = True                                                        # :317-318
            elif i % self.check_time == 0 and not self.preburner and \
                    (self.best_loss < self.int_loss or i > 300):
                self.stale_save = self.best_loss                                   # :319-321
        if not stop:
            if self.preburner and ((loss < 0.25 * self.best_loss) or i > self.preburn_time):   # :325-327
                self.preburner = False
        return stop
