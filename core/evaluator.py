"""Evaluators (host-side metrics; interface of the reference's core/evaluator.py)."""
import numpy as np


class BaseEvaluator(object):

    @classmethod
    def evaluate(cls, predictions, targets):
        raise NotImplementedError("Must specify evaluator.")


class AccEvaluator(BaseEvaluator):

    @classmethod
    def evaluate(cls, predictions, targets):
        predictions, targets = np.asarray(predictions), np.asarray(targets)
        total_num = len(predictions)
        hit_num = int(np.sum(predictions == targets))
        return {"total_num": total_num, "hit_num": hit_num, "accuracy": 1.0 * hit_num / total_num}


class _Unimplemented(BaseEvaluator):
    """Precision / Recall / F1 / ROC / R2 are empty stubs upstream (evaluator.py:26-60, 110-114)."""

    @classmethod
    def evaluate(cls, predictions, targets):
        return None


class PrecisionEvaluator(_Unimplemented):
    pass


class RecallEvaluator(_Unimplemented):
    pass


class F1Evaluator(_Unimplemented):
    pass


class ROCEvaluator(_Unimplemented):
    pass


class R2Evaluator(_Unimplemented):
    pass


class EVEvaluator(BaseEvaluator):
    """Explained variance: 1 - Var[y - pred] / Var[y], averaged over outputs with Var[y] != 0."""

    @classmethod
    def evaluate(cls, predictions, targets):
        predictions, targets = np.asarray(predictions), np.asarray(targets)
        assert predictions.shape == targets.shape
        axis = None if predictions.ndim == 1 else 0
        diff_var = np.atleast_1d(np.var(targets - predictions, axis=axis))
        target_var = np.atleast_1d(np.var(targets, axis=axis))
        keep = np.where(target_var != 0)[0]
        return {"mean_ev": np.mean(1.0 - diff_var[keep] / target_var[keep])}


def _per_sample(err, ndim):
    if ndim == 1:
        return np.mean(err)
    if ndim == 2:
        return np.mean(np.sum(err, axis=1))
    raise ValueError("predision supposes to have 1 or 2 dim.")


class MSEEvaluator(BaseEvaluator):

    @classmethod
    def evaluate(cls, predictions, targets):
        predictions, targets = np.asarray(predictions), np.asarray(targets)
        assert predictions.shape == targets.shape
        return {"mse": _per_sample(np.square(predictions - targets), predictions.ndim)}


class MAEEvaluator(BaseEvaluator):

    @classmethod
    def evaluate(cls, predictions, targets):
        predictions, targets = np.asarray(predictions), np.asarray(targets)
        assert predictions.shape == targets.shape
        # the reference reports this under the key "mse" as well (evaluator.py:106)
        return {"mse": _per_sample(np.abs(predictions - targets), predictions.ndim)}
