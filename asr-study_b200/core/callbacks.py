"""Training callbacks with the reference's surface (core/callbacks.py:8-56 and the keras.callbacks classes train.py
reaches through ``--lr_schedule``, train.py:165-170).  Host-side bookkeeping of the training loop; no kernels.

``MetaCheckpoint`` saves the model together with the ``meta`` record (training arguments, the list of finished epochs
and every logged metric per epoch: what core/callbacks.py:36-56 writes into the HDF5 ``meta`` group) through
``CTCModel.save``.  NOTE on the reference: its constructor forwards the constants monitor='val_loss',
save_best_only=False to ``ModelCheckpoint`` whatever the caller passed (core/callbacks.py:21-24), so the reference's
``best.h5`` (train.py:158-160: monitor='val_decoder_ler', save_best_only=True, mode='min') is in fact rewritten every
epoch.  Here the arguments are honoured — best.* keeps the epoch with the lowest val_decoder_ler, which is what train.py
asks for and what its final test evaluation expects; ``reference_quirk=True`` reproduces the reference's behaviour.
"""
from __future__ import annotations

import os

import numpy as np


class Callback(object):
    def set_model(self, model):
        self.model = model

    def on_epoch_end(self, epoch, logs=None):
        pass


class MetaCheckpoint(Callback):
    def __init__(self, filepath, monitor="val_loss", verbose=0, save_best_only=False, save_weights_only=False,
                 mode="auto", period=1, training_args=None, meta=None, reference_quirk=False):
        if reference_quirk:
            monitor, save_best_only, mode = "val_loss", False, "auto"
        self.filepath, self.monitor, self.verbose = filepath, monitor, verbose
        self.save_best_only, self.period = bool(save_best_only), int(period)
        self.meta = dict(meta) if meta else {"epochs": []}
        self.meta["epochs"] = list(self.meta.get("epochs", []))
        if training_args is not None:
            self.meta["training_args"] = dict(vars(training_args)) if not isinstance(training_args, dict) else dict(training_args)
        if mode not in ("auto", "min", "max"):
            mode = "auto"
        if mode == "auto":
            mode = "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"
        self.better = (lambda a, b: a < b) if mode == "min" else (lambda a, b: a > b)
        self.best = np.inf if mode == "min" else -np.inf
        self.epochs_since_last_save = 0

    def on_epoch_end(self, epoch, logs=None):
        logs = logs or {}
        self.meta["epochs"].append(int(epoch))
        for k, v in logs.items():
            self.meta.setdefault(k, []).append(float(v))
        self.epochs_since_last_save += 1
        if self.epochs_since_last_save < self.period:
            return
        if self.save_best_only:
            cur = logs.get(self.monitor)
            if cur is None or not self.better(cur, self.best):
                return
            self.best = cur
        self.epochs_since_last_save = 0
        if int(os.environ.get("RANK", "0")) == 0:            # data parallel: the replicas are identical, one writer
            self.model.save(self.filepath.format(epoch=epoch, **logs), meta=self.meta)


class ReduceLROnPlateau(Callback):
    """keras.callbacks.ReduceLROnPlateau (Keras 1.2.2 defaults)."""

    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", epsilon=1e-4, cooldown=0,
                 min_lr=0):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, float(factor), int(patience), verbose
        self.epsilon, self.cooldown, self.min_lr = float(epsilon), int(cooldown), float(min_lr)
        if mode == "max" or (mode == "auto" and "acc" in monitor):
            self.better, self.best = (lambda a, b: a > b + self.epsilon), -np.inf
        else:
            self.better, self.best = (lambda a, b: a < b - self.epsilon), np.inf
        self.wait, self.cooldown_counter = 0, 0

    def on_epoch_end(self, epoch, logs=None):
        cur = (logs or {}).get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        if self.better(cur, self.best):
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            if self.wait >= self.patience:
                old = float(self.model.optimizer.lr)
                if old > self.min_lr + 1e-4 * self.min_lr:
                    self.model.optimizer.lr = max(old * self.factor, self.min_lr)
                    if self.verbose:
                        print("Epoch %05d: reducing learning rate to %s." % (epoch, self.model.optimizer.lr))
                    self.cooldown_counter, self.wait = self.cooldown, 0
            self.wait += 1


class LearningRateScheduler(Callback):
    """keras.callbacks.LearningRateScheduler: schedule(epoch) -> lr, applied at the START of each epoch in Keras; this
    loop has no epoch-begin hook, so the value for epoch e + 1 is set at the end of epoch e (and for the first epoch by
    train.py before fitting)."""

    def __init__(self, schedule):
        self.schedule = schedule

    def on_epoch_end(self, epoch, logs=None):
        self.model.optimizer.lr = float(self.schedule(epoch + 1))


LR_SCHEDULES = {"reducelronplateau": ReduceLROnPlateau, "learningratescheduler": LearningRateScheduler}
