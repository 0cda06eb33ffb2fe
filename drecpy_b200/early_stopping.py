"""MaxValidationValueRule with the compute / stop_training protocol fit() drives
(DRecPy/Recommender/EarlyStopping/early_stopping_rule_abc.py:22-66, max_validation_value_rule.py:4-30).
Any object with the same two methods (e.g. the reference's own rules) can be passed to fit()."""
from .recommender import InvalidEpochValidationResultsException


class MaxValidationValueRule:
    def __init__(self, validation_metric, **kwds):
        self.validation_metric = validation_metric
        self.required_validation_metrics = [validation_metric]

    def compute(self, epoch_losses, epoch_validation_results, called_epochs_validation_results, **kwds):
        if not isinstance(epoch_validation_results, dict) or len(epoch_validation_results) == 0:
            raise InvalidEpochValidationResultsException('Epoch callback results must be a non-empty dict.')
        if '@' in self.validation_metric:
            names = {m: m for m in epoch_validation_results}
        else:
            names = {m.split('@')[0]: m for m in epoch_validation_results}
        if self.validation_metric not in names or len(epoch_validation_results[names[self.validation_metric]]) == 0:
            raise InvalidEpochValidationResultsException(
                f'No matching epoch callback metric with the required validation metrics. Expected: '
                f'{self.required_validation_metrics}, found: {names.keys()}.')
        vals = epoch_validation_results[names[self.validation_metric]]
        best_idx, best = 0, vals[0]
        for idx, v in enumerate(vals):
            if v > best:
                best, best_idx = v, idx
        return called_epochs_validation_results[best_idx]

    def stop_training(self, current_epoch, best_computed_epoch, target_epoch, **kwds):
        return False
