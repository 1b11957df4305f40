"""nuspacesim_b200: the per-event optical Monte-Carlo hot path of nuSpaceSim on B200.

Drop-in for ``nuspacesim.compute`` / ``RegionGeom`` / ``Taus`` / ``EAS`` (optical branch);
every stage runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/nss_b200.h`` (``libnss_b200.so``).  There is no CPU fallback.
"""
from .config import NssConfig, config_from_toml, create_toml  # noqa: F401
from .compute import compute, run_acceptance, AcceptanceRunner  # noqa: F401
from .simulation.geometry.region_geometry import RegionGeom  # noqa: F401
from .simulation.taus.taus import Taus  # noqa: F401
from .simulation.eas_optical.eas import EAS  # noqa: F401
from .simulation.eas_optical.cphotang import CphotAng  # noqa: F401
from .simulation.spectra.spectra import Spectra  # noqa: F401
from .simulation.atmosphere.clouds import CloudTopHeight  # noqa: F401
from .utils.grid import NssGrid  # noqa: F401
from . import results_table  # noqa: F401

__all__ = [
    "NssConfig", "config_from_toml", "create_toml", "compute", "run_acceptance",
    "AcceptanceRunner", "RegionGeom", "Taus", "EAS", "CphotAng", "Spectra", "CloudTopHeight",
    "NssGrid", "results_table",
]
