"""Configuration objects with the attribute layout of the reference's ``NssConfig``.

Reference: src/nuspacesim/config.py:72-393 (pydantic models + astropy unit parsing).
This mirror keeps the same nesting, field names, defaults, units (km / rad / m2 / MHz)
and TOML layout so ``config_from_toml("sample_input_file.toml")`` works unchanged, but
uses plain dataclasses and a small unit parser (astropy is not a dependency here).
Only the fields the optical hot path reads are interpreted; the rest are carried.
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass, field, fields, is_dataclass
from typing import Any, Optional, Union

try:  # Python >= 3.11
    import tomllib
except ImportError:  # pragma: no cover
    import tomli as tomllib

__all__ = ["NssConfig", "Detector", "Simulation", "config_from_toml", "create_toml", "parse_units"]

_UNIT_SCALE = {
    # length -> km
    ("km", "km"): 1.0, ("m", "km"): 1e-3, ("cm", "km"): 1e-5,
    # angle -> rad
    ("rad", "rad"): 1.0, ("deg", "rad"): math.pi / 180.0,
    # area -> m2
    ("m2", "m2"): 1.0, ("cm2", "m2"): 1e-4,
    # frequency -> MHz
    ("MHz", "MHz"): 1.0, ("Hz", "MHz"): 1e-6, ("kHz", "MHz"): 1e-3, ("GHz", "MHz"): 1e3,
    ("dB", "dB"): 1.0,
}


def parse_units(value: Union[float, int, str], unit: str) -> float:
    """'525.0 km' -> 525.0 (in ``unit``); bare numbers are already in ``unit``.

    Reference: config.py:72-76 (astropy ``Quantity(value).to(unit).value``).
    """
    if isinstance(value, str):
        parts = value.split()
        if len(parts) == 1:
            return float(parts[0])
        if len(parts) != 2 or (parts[1], unit) not in _UNIT_SCALE:
            raise ValueError(f"cannot convert {value!r} to {unit}")
        return float(parts[0]) * _UNIT_SCALE[(parts[1], unit)]
    return float(value)


def _build(cls, data):
    """Construct dataclass ``cls`` from a (possibly nested) dict, or pass through."""
    if data is None or isinstance(data, cls):
        return data
    if not isinstance(data, dict):
        raise TypeError(f"{cls.__name__}: expected a mapping, got {type(data).__name__}")
    known = {f.name for f in fields(cls)}
    unknown = set(data) - known
    if unknown:
        raise ValueError(f"{cls.__name__}: unknown field(s) {sorted(unknown)}")
    return cls(**data)


@dataclass
class InitialPos:
    altitude: float = 525.0  # km
    latitude: float = 0.0  # rad
    longitude: float = 0.0  # rad

    def __post_init__(self):
        self.altitude = parse_units(self.altitude, "km")
        self.latitude = parse_units(self.latitude, "rad")
        self.longitude = parse_units(self.longitude, "rad")


@dataclass
class SunMoon:
    sun_moon_cuts: bool = True
    sun_alt_cut: float = math.radians(-18.0)
    moon_alt_cut: float = math.radians(0.0)
    moon_min_phase_angle_cut: float = math.radians(150.0)

    def __post_init__(self):
        for n in ("sun_alt_cut", "moon_alt_cut", "moon_min_phase_angle_cut"):
            setattr(self, n, parse_units(getattr(self, n), "rad"))


@dataclass
class Optical:
    enable: bool = True
    telescope_effective_area: float = 2.5  # m2
    quantum_efficiency: float = 0.2
    photo_electron_threshold: float = 10

    def __post_init__(self):
        self.telescope_effective_area = parse_units(self.telescope_effective_area, "m2")
        self.quantum_efficiency = float(self.quantum_efficiency)
        self.photo_electron_threshold = float(self.photo_electron_threshold)


@dataclass
class Radio:
    enable: bool = True
    low_frequency: float = 30.0
    high_frequency: float = 300.0
    snr_threshold: float = 5.0
    nantennas: int = 10
    gain: float = 1.8

    def __post_init__(self):
        self.low_frequency = parse_units(self.low_frequency, "MHz")
        self.high_frequency = parse_units(self.high_frequency, "MHz")
        self.gain = parse_units(self.gain, "dB")
        if self.high_frequency <= self.low_frequency:
            raise ValueError("High frequency must be greater than low frequency")


@dataclass
class Detector:
    name: str = "Default Name"
    initial_position: InitialPos = field(default_factory=InitialPos)
    sun_moon: Optional[SunMoon] = field(default_factory=SunMoon)
    optical: Optional[Optical] = field(default_factory=Optical)
    radio: Optional[Radio] = field(default_factory=Radio)

    InitialPos = InitialPos
    SunMoon = SunMoon
    Optical = Optical
    Radio = Radio

    def __post_init__(self):
        self.initial_position = _build(InitialPos, self.initial_position)
        self.sun_moon = _build(SunMoon, self.sun_moon)
        self.optical = _build(Optical, self.optical)
        self.radio = _build(Radio, self.radio)


@dataclass
class Ionosphere:
    enable: bool = True
    total_electron_content: float = 10.0
    total_electron_error: float = 0.1


@dataclass
class NuPyPropShower:
    id: str = "nupyprop"
    etau_frac: float = 0.5
    table_version: str = "3"

    def __post_init__(self):
        if self.id not in ("nupyprop", "nuleptonsim", "nupyprop_bdhm"):
            raise ValueError(f"unknown tau_shower id {self.id!r}")
        self.table_version = str(self.table_version)
        self.etau_frac = float(self.etau_frac)


@dataclass
class MonoSpectrum:
    id: str = "monospectrum"
    log_nu_energy: float = 8.0


@dataclass
class PowerSpectrum:
    id: str = "powerspectrum"
    index: float = 2.0
    lower_bound: float = 6.0
    upper_bound: float = 12.0


@dataclass
class NoCloud:
    id: str = "no_cloud"


@dataclass
class MonoCloud:
    id: str = "monocloud"
    altitude: float = float("-inf")


@dataclass
class PressureMapCloud:
    id: str = "pressure_map"
    month: int = 1
    version: Union[str, int] = 0


@dataclass
class TargetOfOpportunity:
    source_RA: float = 0.0
    source_DEC: float = 0.0
    source_date: str = "2022-06-02T01:00:00"
    source_date_format: str = "isot"
    source_obst: float = 86400

    def __post_init__(self):
        self.source_RA = parse_units(self.source_RA, "rad")
        self.source_DEC = parse_units(self.source_DEC, "rad")


def _by_id(data, options, what):
    if data is None or is_dataclass(data):
        return data
    if not isinstance(data, dict) or "id" not in data:
        raise ValueError(f"{what}: mapping with an 'id' key expected")
    for cls in options:
        if cls().id == data["id"]:
            return _build(cls, data)
    raise ValueError(f"{what}: unknown id {data['id']!r}")


@dataclass
class Simulation:
    mode: str = "Diffuse"
    thrown_events: int = 1000
    max_cherenkov_angle: float = math.radians(3)
    max_azimuth_angle: float = math.radians(360)
    angle_from_limb: float = math.radians(7)
    eas_long_profile: str = "Greisen"
    cherenkov_light_engine: str = "nuspacesim"
    ionosphere: Optional[Ionosphere] = field(default_factory=Ionosphere)
    tau_shower: NuPyPropShower = field(default_factory=NuPyPropShower)
    spectrum: Union[MonoSpectrum, PowerSpectrum] = field(default_factory=MonoSpectrum)
    cloud_model: Union[NoCloud, MonoCloud, PressureMapCloud] = field(default_factory=NoCloud)
    target: Optional[TargetOfOpportunity] = field(default_factory=TargetOfOpportunity)

    Ionosphere = Ionosphere
    NuPyPropShower = NuPyPropShower
    MonoSpectrum = MonoSpectrum
    PowerSpectrum = PowerSpectrum
    NoCloud = NoCloud
    MonoCloud = MonoCloud
    PressureMapCloud = PressureMapCloud
    TargetOfOpportunity = TargetOfOpportunity

    def __post_init__(self):
        if self.mode not in ("Diffuse", "Target"):
            raise ValueError(f"mode must be 'Diffuse' or 'Target', got {self.mode!r}")
        self.thrown_events = int(self.thrown_events)
        for n in ("max_cherenkov_angle", "max_azimuth_angle", "angle_from_limb"):
            setattr(self, n, parse_units(getattr(self, n), "rad"))
        if self.eas_long_profile == "Default":
            self.eas_long_profile = "Greisen"
        if self.eas_long_profile not in (
            "Greisen", "Gaisser-Hillas Parameterized", "Gaisser-Hillas Fluctuated"
        ):
            raise ValueError(f"unknown eas_long_profile {self.eas_long_profile!r}")
        if self.cherenkov_light_engine == "Default":
            self.cherenkov_light_engine = "nuspacesim"
        if self.cherenkov_light_engine != "nuspacesim":
            raise ValueError(f"unknown cherenkov_light_engine {self.cherenkov_light_engine!r}")
        self.ionosphere = _build(Ionosphere, self.ionosphere)
        self.tau_shower = _build(NuPyPropShower, self.tau_shower)
        self.spectrum = _by_id(self.spectrum, (MonoSpectrum, PowerSpectrum), "spectrum")
        self.cloud_model = _by_id(
            self.cloud_model, (NoCloud, MonoCloud, PressureMapCloud), "cloud_model"
        )
        self.target = _build(TargetOfOpportunity, self.target)


@dataclass
class NssConfig:
    """Container of the detector and simulation parameters (config.py:370-382)."""

    title: str = "NuSpaceSim"
    detector: Detector = field(default_factory=Detector)
    simulation: Simulation = field(default_factory=Simulation)

    def __post_init__(self):
        self.detector = _build(Detector, self.detector)
        self.simulation = _build(Simulation, self.simulation)

    def model_dump(self) -> dict:
        """Nested dict with unit strings, like the reference's serializers."""
        d = asdict(self)
        ip = d["detector"]["initial_position"]
        ip["altitude"] = f"{ip['altitude']} km"
        ip["latitude"] = f"{math.degrees(ip['latitude'])} deg"
        ip["longitude"] = f"{math.degrees(ip['longitude'])} deg"
        if d["detector"].get("sun_moon"):
            sm = d["detector"]["sun_moon"]
            for n in ("sun_alt_cut", "moon_alt_cut", "moon_min_phase_angle_cut"):
                sm[n] = f"{math.degrees(sm[n])} deg"
        if d["detector"].get("optical"):
            o = d["detector"]["optical"]
            o["telescope_effective_area"] = f"{o['telescope_effective_area']} m2"
        if d["detector"].get("radio"):
            r = d["detector"]["radio"]
            r["low_frequency"] = f"{r['low_frequency']} MHz"
            r["high_frequency"] = f"{r['high_frequency']} MHz"
            r["gain"] = f"{r['gain']} dB"
        s = d["simulation"]
        for n in ("max_cherenkov_angle", "max_azimuth_angle", "angle_from_limb"):
            s[n] = f"{math.degrees(s[n])} deg"
        if s.get("target"):
            for n in ("source_RA", "source_DEC"):
                s["target"][n] = f"{math.degrees(s['target'][n])} deg"
        return d


def config_from_toml(filename: str) -> NssConfig:
    """config.py:385-388."""
    with open(filename, "rb") as f:
        return NssConfig(**tomllib.load(f))


def _toml_value(v: Any) -> str:
    if isinstance(v, bool):
        return "true" if v else "false"
    if isinstance(v, (int, float)):
        if isinstance(v, float) and math.isinf(v):
            return "inf" if v > 0 else "-inf"
        return repr(v)
    return '"' + str(v).replace("\\", "\\\\").replace('"', '\\"') + '"'


def _toml_dump(d: dict, prefix: str, out: list) -> None:
    scalars = {k: v for k, v in d.items() if not isinstance(v, dict) and v is not None}
    tables = {k: v for k, v in d.items() if isinstance(v, dict)}
    if prefix and scalars:
        out.append(f"[{prefix}]")
    for k, v in scalars.items():
        out.append(f"{k} = {_toml_value(v)}")
    if scalars:
        out.append("")
    for k, v in tables.items():
        _toml_dump(v, f"{prefix}.{k}" if prefix else k, out)


def create_toml(filename: str, c: NssConfig) -> None:
    """config.py:391-393 (small TOML emitter; tomli_w is not required)."""
    out: list = []
    _toml_dump(c.model_dump(), "", out)
    with open(filename, "w") as f:
        f.write("\n".join(out))
