"""Loss / callback history with the attribute names early-stopping rules read
(DRecPy/Evaluation/loss_tracker.py:4-50; the matplotlib plot of :52-94 is out of scope)."""


class LossTracker:
    def __init__(self):
        self.epoch_losses = []
        self.curr_avg_epoch_loss = 0
        self.epoch_callback_results = {}
        self.called_epochs = []

    def add_epoch_loss(self, loss):
        self.epoch_losses.append(loss)
        self.curr_avg_epoch_loss = self.curr_avg_epoch_loss + (loss - self.curr_avg_epoch_loss) / len(self.epoch_losses)

    def get_epoch_avg_loss(self):
        return self.curr_avg_epoch_loss

    def reset_epoch_losses(self):
        self.epoch_losses = []
        self.curr_avg_epoch_loss = 0

    def add_epoch_callback_result(self, name, result, epoch):
        if name not in self.epoch_callback_results:
            self.epoch_callback_results[name] = []
        self.epoch_callback_results[name].append(result)
        if len(self.called_epochs) == 0 or self.called_epochs[-1] < epoch:
            self.called_epochs.append(epoch)
