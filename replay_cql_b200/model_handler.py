"""``save`` / ``load`` with the reference's directory layout (``replay/model_handler.py:29-92``):

``path/model`` (owned by ``_save_model``), ``path/init_args.json`` (``_init_args`` +
``_model_name``), ``path/dataframes/{fit_users,fit_items,...}`` (parquet), ``path/study`` (joblib).
The class is resolved by name, as the reference does via ``globals()`` (``:69``).
"""
from __future__ import annotations

import json
import os
import shutil
from inspect import getfullargspec
from os.path import exists, join

import joblib
import pandas as pd

from . import models as _models
from .recommender import Recommender


def prepare_dir(path: str) -> None:
    if exists(path):
        shutil.rmtree(path)
    os.makedirs(path)


def save(model: Recommender, path: str) -> None:
    prepare_dir(path)
    model._save_model(join(path, "model"))
    init_args = dict(model._init_args)
    init_args["_model_name"] = str(model)
    with open(join(path, "init_args.json"), "w") as f:
        json.dump(init_args, f)
    df_path = join(path, "dataframes")
    os.makedirs(df_path)
    for name, df in model._dataframes.items():
        df.to_parquet(join(df_path, name), index=False)
    model.fit_users.to_parquet(join(df_path, "fit_users"), index=False)
    model.fit_items.to_parquet(join(df_path, "fit_items"), index=False)
    joblib.dump(model.study, join(path, "study"))


def load(path: str) -> Recommender:
    with open(join(path, "init_args.json")) as f:
        args = json.load(f)
    name = args.pop("_model_name")
    model_class = getattr(_models, name)
    init_names = getfullargspec(model_class.__init__).args
    init_names.remove("self")
    extra = set(args) - set(init_names)
    init_args = {k: args[k] for k in init_names if k in args}
    model = model_class(**init_args)
    for arg in extra:   # reference quirk (model_handler.py:81-82): extras land on a literal attribute `arg`
        model.arg = args[arg]
    df_path = join(path, "dataframes")
    for fname in os.listdir(df_path):
        setattr(model, fname, pd.read_parquet(join(df_path, fname)))
    model._load_model(join(path, "model"))
    model.study = joblib.load(join(path, "study"))
    return model
