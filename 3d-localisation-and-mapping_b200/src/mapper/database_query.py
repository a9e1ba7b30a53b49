"""RTAB-Map pose file -> DataFrame (same class / method / columns as the reference's
``PoseDataExtractor.fetch_data``, ``/root/reference/src/mapper/database_query.py:12-25``).
``plot_pose`` is an Open3D GUI and is out of scope."""
from __future__ import annotations

import pandas as pd


class PoseDataExtractor:
    def __init__(self, pose_path):
        self.pose_path = pose_path

    def fetch_data(self):
        df = pd.read_csv(self.pose_path, sep=" ", skiprows=1, header=None)
        df.columns = ["timestamp", "tx", "ty", "tz", "qx", "qy", "qz", "qw", "id"]
        df["timestamp"] = pd.to_datetime(df["timestamp"], unit="s")
        df = df.drop(["id"], axis=1)
        return df

    def plot_pose(self, df):
        raise NotImplementedError("Open3D pose plot is outside the B200 lift's scope (SURVEY.md section 2)")
