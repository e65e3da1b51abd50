"""On-disk side of the predict task: the parquet format written by `tasks/predict.py:52-64` through
`utils/miscellaneous.py:12-41` (`ResultSaver`), i.e. one row per window with the columns `labels, ep_vars, al_vars,
preds, stds` produced by `BNN.predict_step` / `HNN.predict_step`, and the reader `results/predictions.py:31-40` expects.
Host-side glue only: the numbers come from brl_predict_moments(_host)."""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Iterable, List, Union

import numpy as np
import pandas as pd

PREDICTION_COLUMNS = ["labels", "ep_vars", "al_vars", "preds", "stds"]


class ResultSaver:
    """utils/miscellaneous.py:12-41 -- same constructor, `save` / `load` / `append`, same assertions."""

    def __init__(self, path: Union[Path, str], filename: str) -> None:
        self.path = Path(path)
        self.path.mkdir(exist_ok=True)
        self.file_path = Path(self.path, filename)

    def save(self, df: Union[pd.DataFrame, dict]) -> None:
        if isinstance(df, dict):
            df = pd.DataFrame(df)
        assert isinstance(df, pd.DataFrame), f"{type(df)} is not a dataframe"
        df.to_parquet(self.file_path)

    def load(self) -> pd.DataFrame:
        return pd.read_parquet(self.file_path)

    def append(self, series: Union[List[pd.Series], Dict[str, np.ndarray]]) -> None:
        if isinstance(series, list):
            series = pd.concat(series, axis=1)
        if isinstance(series, dict):
            series = pd.DataFrame(series)
        df = pd.concat([self.load(), series], axis=1)
        assert isinstance(df, pd.DataFrame), f"{type(df)} is not a dataframe"
        assert int(np.asarray(df.isna().sum()).sum()) == 0, "NaNs introduced in results dataframe"
        self.save(df)


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
    ResultSaver(out_dir, filename).save(frame)
    return frame
