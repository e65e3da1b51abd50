"""On-disk side of the predict task: the parquet format written by `tasks/predict.py:52-64`, i.e. one row per
window with the columns `labels, ep_vars, al_vars, preds, stds` produced by `BNN.predict_step` / `HNN.predict_step`, which is
what the reader `results/predictions.py:31-40` expects.  Host-side glue only: the numbers come from
brl_predict_moments(_host); the reference's own `ResultSaver` (file IO, out of scope) keeps working on these frames."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, Union

import numpy as np
import pandas as pd

PREDICTION_COLUMNS = ["labels", "ep_vars", "al_vars", "preds", "stds"]


def predictions_to_frame(predictions: Iterable[Dict[str, np.ndarray]]) -> pd.DataFrame:
    """tasks/predict.py:57-60: the list of per-batch `predict_step` dicts -> one row per window (from_records + explode).
    Built by concatenation (same frame, without the object-dtype detour of `explode`)."""
    predictions = list(predictions)
    if not predictions:
        return pd.DataFrame({c: np.zeros(0, dtype=np.float32) for c in PREDICTION_COLUMNS})
    cols = list(predictions[0].keys())
    return pd.DataFrame({c: np.concatenate([np.atleast_1d(np.asarray(p[c])) for p in predictions]) for c in cols})


def write_predictions(model, batches: Iterable, out_dir: Union[Path, str], filename: str) -> pd.DataFrame:
    """The body of tasks/predict.py:52-64 for one (checkpointed) model and one subset: run `predict_step` over the batches
    of the predict dataloader (order preserved: `shuffle=False` is load-bearing, data/ncmapss/dataset.py:123-139) and
    save `<method>_<run>_<subset>.parquet`."""
    if hasattr(model, "on_predict_start"):
        model.on_predict_start()
    frame = predictions_to_frame(model.predict_step(batch, i) for i, batch in enumerate(batches))
    out_dir = Path(out_dir)
    out_dir.mkdir(parents=True, exist_ok=True)
    frame.to_parquet(out_dir / filename)
    return frame
