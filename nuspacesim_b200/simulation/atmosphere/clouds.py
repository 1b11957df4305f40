"""CloudTopHeight: cloud-top altitude callable handed to the EAS stage.

Mirror of ``nuspacesim.simulation.atmosphere.clouds.CloudTopHeight``
(src/nuspacesim/simulation/atmosphere/clouds.py:54-82) for the NoCloud and MonoCloud
models, which are constants and travel to the kernel as one scalar.  PressureMapCloud
(FITS maps) is outside the accelerated path.
"""
from __future__ import annotations

import numpy as np

__all__ = ["CloudTopHeight"]


class CloudTopHeight:
    def __init__(self, config):
        self.cloud_model = config.simulation.cloud_model
        cid = getattr(self.cloud_model, "id", "no_cloud") if self.cloud_model else "no_cloud"
        if cid == "no_cloud":
            self._alt = 0.0
        elif cid == "monocloud":
            self._alt = float(self.cloud_model.altitude)
        else:
            raise NotImplementedError(f"cloud model {cid!r} is not implemented")

    def __call__(self, *args, **kwargs) -> np.single:
        return np.single(self._alt)
