"""``python -m nuspacesim_b200 run cfg.toml [count] ...``: the reference's ``run`` command
(src/nuspacesim/apps/run.py:134-210, cli.py:54-61) over the B200 hot path.

Same positional arguments and the options that concern the accelerated path (``-o``, ``-w``,
``-n``, ``--monospectrum``, ``--powerspectrum``, ``--nocloud``, ``--monocloud``).  Plot options
are accepted and ignored (plotting is outside the path); ``--pressuremapcloud`` raises.
Extras: ``--seed`` (Philox key; the reference is unseeded) and ``--acceptance-only`` (column-free
mode: only the four Monte-Carlo integral scalars, for very large counts).
"""
from __future__ import annotations

import click

from ..config import (MonoCloud, MonoSpectrum, NoCloud, NssConfig, PowerSpectrum,
                      config_from_toml, create_toml)
from ..results_table import output_filename


@click.group()
def cli():
    pass


@cli.command()
@click.option("-o", "--output", type=click.Path(exists=False), default=None, help="Output file.")
@click.option("-p", "--plot", multiple=True, default=[], help="Ignored (no plotting here).")
@click.option("-P", "--plotconfig", default=None, help="Ignored (no plotting here).")
@click.option("--plotall", is_flag=True, help="Ignored (no plotting here).")
@click.option("-w", "--write-stages", is_flag=True,
              help="Write intermediate values after each simulation stage.")
@click.option("-n", "--no-result-file", is_flag=True,
              help="Do not save the results to an output file.")
@click.option("--monospectrum", type=float, default=None,
              help="Mono Energetic Spectrum Log Energy.")
@click.option("--powerspectrum", nargs=3, type=click.Tuple([float, float, float]), default=None,
              help="Power Spectrum index, lower_bound, upper_bound.")
@click.option("--nocloud", is_flag=True, default=None, help="No Cloud Model. [Default]")
@click.option("--monocloud", type=float, default=None,
              help="Uniform (mono) Height Cloud Model (km).")
@click.option("--pressuremapcloud", default=None, help="Not supported on the accelerated path.")
@click.option("--seed", type=int, default=0, help="Philox key of the run.")
@click.option("--acceptance-only", is_flag=True,
              help="Column-free mode: print the Monte-Carlo integral only.")
@click.argument("config_file", type=click.Path(exists=True, dir_okay=False, readable=True))
@click.argument("count", type=float, default=0.0)
def run(config_file, count, no_result_file, output, plot, plotconfig, plotall, write_stages,
        monospectrum, powerspectrum, nocloud, monocloud, pressuremapcloud, seed,
        acceptance_only):
    """Perform the full nuspacesim simulation (optical branch) on the GPU."""
    from ..compute import compute, run_acceptance

    config = config_from_toml(config_file)
    config.simulation.thrown_events = int(
        config.simulation.thrown_events if count == 0.0 else count)
    # apps/utils.py:8-31 (parse_spectra_options / parse_cloud_options)
    if monospectrum is not None and powerspectrum:
        raise click.UsageError("Only one of --monospectrum, --powerspectrum may be used.")
    if monospectrum is not None:
        config.simulation.spectrum = MonoSpectrum(log_nu_energy=monospectrum)
    if powerspectrum:
        config.simulation.spectrum = PowerSpectrum(
            index=powerspectrum[0], lower_bound=powerspectrum[1], upper_bound=powerspectrum[2])
    if pressuremapcloud is not None:
        raise click.UsageError("--pressuremapcloud is outside the accelerated path")
    if sum(x is not None for x in (nocloud, monocloud)) > 1:
        raise click.UsageError("Only one cloud model may be selected.")
    if nocloud:
        config.simulation.cloud_model = NoCloud()
    if monocloud is not None:
        config.simulation.cloud_model = MonoCloud(altitude=monocloud)

    if acceptance_only:
        mcint, geo, npass, unc = run_acceptance(config, config.simulation.thrown_events,
                                                seed=seed)
        click.echo(f"Monte Carlo Integral [Optical]: {mcint}")
        click.echo(f"Monte Carlo Integral, GEO Only [Optical]: {geo}")
        click.echo(f"Number of Passing Events [Optical]: {npass}")
        click.echo(f"Stat uncert of MC Integral [Optical]: {unc}")
        return
    output = output_filename(output)
    simulation = compute(config, verbose=True, to_plot=[], output_file=output,
                         write_stages=write_stages, seed=seed)
    if not no_result_file:
        simulation.write(output, format="fits", overwrite=True)


@cli.command(name="create-config")
@click.option("-n", "--numtrajs", type=float, default=100, help="number of thrown events")
@click.option("--monospectrum", type=float, default=None)
@click.option("--powerspectrum", nargs=3, type=click.Tuple([float, float, float]), default=None)
@click.argument("filename", type=click.Path(exists=False, writable=True))
def create_config(filename, numtrajs, monospectrum, powerspectrum):
    """Generate a configuration file from the default configurations
    (src/nuspacesim/apps/create_config.py)."""
    c = NssConfig()
    c.simulation.thrown_events = int(numtrajs)
    if monospectrum is not None:
        c.simulation.spectrum = MonoSpectrum(log_nu_energy=monospectrum)
    if powerspectrum:
        c.simulation.spectrum = PowerSpectrum(index=powerspectrum[0],
                                              lower_bound=powerspectrum[1],
                                              upper_bound=powerspectrum[2])
    create_toml(filename, c)


if __name__ == "__main__":
    cli()
