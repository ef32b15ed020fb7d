"""SUMMARY -- host-side mirror of the reference's src/summary.py.

The reference logs to TensorBoard through TensorFlow summary ops; here the same call surface
(`Summary(options, session, summary_path)` with `add`, `add_to_eval_summary`,
`add_to_training_summary`, `add_to_overlap_summary`, `add_to_pixel_missclassification_summary`,
`add_to_eval_patch_summary`, `img_to_label_patches`, `get_prediction_metrics`, `flush`) writes

  <summary_path>/scalars.jsonl          one {"tag", "value", "step"} record per scalar
  <summary_path>/images/<tag>_<step>_<i>.png   what the reference sends as image summaries

and keeps every scalar in memory (`Summary.scalars`) for the caller.  Patch labels come from the
`rsu_patch_vote` kernel (images.patch_labels); the streaming accuracy / recall / precision / F1 are
the host-side restatement of tf.metrics.* (summary.py:141-147): running counts that every update
adds to and that are zeroed once per epoch (`reset`, the `tf.local_variables_initializer().run()`
of tf_aerial_images.py:428).  The evaluation and the training summaries own separate counters,
like the separate tf.metrics calls at summary.py:44 and :66 do.
"""
import json
import os

import numpy as np

from . import images
from .constants import IMG_PATCH_SIZE


class StreamingMetrics:
    """tf.metrics.accuracy / recall / precision as the reference combines them
    (summary.py:141-147): counts accumulate over update() calls until reset()."""

    def __init__(self):
        self.reset()

    def reset(self):
        self.total = 0.0
        self.correct = 0.0
        self.true_positives = 0.0
        self.false_negatives = 0.0
        self.false_positives = 0.0

    def update(self, labels, predictions, padded_zeros=0):
        """labels / predictions: integer arrays of equal size (non-zero = road).  padded_zeros:
        extra (0, 0) pairs the reference appends through ndarray.resize (summary.py:134-139) --
        they count as correct for the accuracy and touch nothing else.  Returns the values of
        the four update ops: (accuracy, recall, precision, f1_score)."""
        t = np.asarray(labels).reshape(-1) != 0
        p = np.asarray(predictions).reshape(-1) != 0
        assert t.shape == p.shape
        self.total += t.size + padded_zeros
        self.correct += float(np.sum(t == p)) + padded_zeros
        self.true_positives += float(np.sum(p & t))
        self.false_negatives += float(np.sum(~p & t))
        self.false_positives += float(np.sum(p & ~t))
        return self.result()

    def result(self):
        # tf.metrics divides with a guard that yields 0 for an empty denominator
        accuracy = self.correct / self.total if self.total > 0 else 0.0
        pos = self.true_positives + self.false_negatives
        recall = self.true_positives / pos if pos > 0 else 0.0
        claimed = self.true_positives + self.false_positives
        precision = self.true_positives / claimed if claimed > 0 else 0.0
        # 2 / (1 / recall + 1 / precision): TensorFlow evaluates 1/0 = inf and 2/inf = 0
        f1_score = 2.0 / (1.0 / recall + 1.0 / precision) if recall > 0 and precision > 0 else 0.0
        return accuracy, recall, precision, f1_score


class Summary:
    """Handle the run's summaries (summary.py:7-147) without TensorFlow."""

    def __init__(self, options, session, summary_path, write=True):
        self._options = options
        self._session = session  # signature parity, unused
        self._path = summary_path
        self._write = bool(write) and summary_path is not None
        self._file = None
        self.scalars = []  # (tag, value, step)
        self.eval_metrics = StreamingMetrics()
        self.train_metrics = StreamingMetrics()

    # ------------------------------------------------------------------ plumbing
    def _open(self):
        if self._file is None and self._write:
            os.makedirs(self._path, exist_ok=True)
            self._file = open(os.path.join(self._path, "scalars.jsonl"), "a")
        return self._file

    def _scalar(self, tag, value, step):
        value, step = float(value), int(step)
        self.scalars.append((tag, value, step))
        f = self._open()
        if f is not None:
            f.write(json.dumps({"tag": tag, "value": value, "step": step}) + "\n")

    def _images(self, tag, arrays, step, greyscale=False):
        if not self._write:
            return
        arrays = np.asarray(arrays)
        images.save_all(arrays, os.path.join(self._path, "images"),
                        "%s_%06d_{:03d}.png" % (tag, int(step)), greyscale=greyscale)

    def flush(self):
        if self._file is not None:
            self._file.flush()

    def reset(self):
        """tf.local_variables_initializer().run() (tf_aerial_images.py:428): zero the streaming
        counters at the start of every epoch."""
        self.eval_metrics.reset()
        self.train_metrics.reset()

    def last(self, tag):
        for t, v, s in reversed(self.scalars):
            if t == tag:
                return v
        return None

    # ------------------------------------------------------------------ reference surface
    def add(self, scalars, global_step=None):
        """The merged summary_op of tf_aerial_images.py:163-164: {"loss", "learning_rate"}."""
        for key, value in scalars.items():
            self._scalar(key, value, global_step)

    def img_to_label_patches(self, img, patch_size=IMG_PATCH_SIZE):
        """Patch labels of a batch of masks (summary.py:134-139): mean over each patch_size cell
        > FOREGROUND_THRESHOLD.  Returns (labels int64 [n_patches], zeros) where `zeros` is the
        number of zero entries the reference's ndarray.resize((n, p, p)) appends."""
        img = np.asarray(img)
        if img.ndim == 4:
            img = img.squeeze(-1)
        labels = images.patch_labels(img, patch_size).reshape(-1)
        return labels, labels.size * (patch_size * patch_size - 1)

    def get_prediction_metrics(self, labels, predictions, metrics=None, padded_zeros=0):
        m = metrics if metrics is not None else StreamingMetrics()
        return m.update(labels, predictions, padded_zeros)

    def _metric_scalars(self, prefix, metrics, pred_masks, true_masks, step):
        pred, zeros = self.img_to_label_patches(pred_masks)
        true, _ = self.img_to_label_patches(true_masks)
        acc, rec, prec, f1 = self.get_prediction_metrics(true, pred, metrics, zeros)
        for tag, v in (("accuracy", acc), ("recall", rec), ("precision", prec), ("f1_score", f1)):
            self._scalar("%s %s" % (prefix, tag), v, step)
        return acc, rec, prec, f1

    def add_to_eval_summary(self, masks, overlays, labels, global_step):
        """summary.py:104-119: masks / overlays as images, streaming patch metrics of the first
        num_eval_images predictions against their ground truth."""
        opts = self._options
        out = self._metric_scalars("eval", self.eval_metrics, masks,
                                   np.asarray(labels)[:opts.num_eval_images], global_step)
        m = np.asarray(masks)
        self._images("eval_masks", m if m.ndim == 3 else m.squeeze(-1), global_step, greyscale=True)
        self._images("eval_images", overlays, global_step)
        return out

    def add_to_training_summary(self, predictions, labels, global_step):
        """summary.py:121-132: streaming patch metrics of the whole training set."""
        return self._metric_scalars("train", self.train_metrics, predictions, labels, global_step)

    def add_to_overlap_summary(self, true_labels, predicted_labels, global_step):
        """summary.py:78-86: prediction (red) over ground truth (green)."""
        self._images("groundtruth_vs_prediction",
                     images.overlap_pred_true(np.asarray(predicted_labels), np.asarray(true_labels)),
                     global_step)

    def add_to_eval_patch_summary(self, labels):
        """summary.py:88-97: the ground truth of the evaluation images, once per run."""
        opts = self._options
        self._images("eval_groundtruth", np.asarray(labels)[:opts.num_eval_images], 0, greyscale=True)

    def add_to_pixel_missclassification_summary(self, num_errors, total, global_step):
        """summary.py:99-102: sum |label - probability| per trained patch so far this epoch."""
        self._scalar("misclassification_rate", num_errors / total, global_step)
