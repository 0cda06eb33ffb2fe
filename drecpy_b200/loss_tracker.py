"""History of per-epoch losses and epoch-callback results kept by fit().

The attribute and method names are an interface, not a choice: the reference's early-stopping rules and user
callbacks read `epoch_losses`, `epoch_callback_results` and `called_epochs` off the model's `_loss_tracker`
(DRecPy/Evaluation/loss_tracker.py:4-50, consumed at recommender_abc.py:224-232), so an object with exactly this
surface is what fit() must fill.  Any object with the same four methods can be passed to fit(loss_tracker=...)
instead (e.g. the reference's own LossTracker when DRecPy is installed next to this package).  The matplotlib
report of loss_tracker.py:52-94 is out of scope.
"""


class LossTracker:
    def __init__(self):
        self.epoch_losses = []
        self.epoch_callback_results = {}
        self.called_epochs = []
        self.curr_avg_epoch_loss = 0

    def add_epoch_loss(self, loss):
        """Append one loss and keep the running mean current (incremental mean, no re-summation)."""
        self.epoch_losses.append(loss)
        n = len(self.epoch_losses)
        self.curr_avg_epoch_loss += (loss - self.curr_avg_epoch_loss) / n

    def get_epoch_avg_loss(self):
        return self.curr_avg_epoch_loss

    def reset_epoch_losses(self):
        self.epoch_losses, self.curr_avg_epoch_loss = [], 0

    def add_epoch_callback_result(self, name, result, epoch):
        """One metric value reported by the epoch callback at `epoch` (epochs are recorded once, in order)."""
        self.epoch_callback_results.setdefault(name, []).append(result)
        if not self.called_epochs or epoch > self.called_epochs[-1]:
            self.called_epochs.append(epoch)
