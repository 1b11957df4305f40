"""Command line driver compatible with ``nuspacesim run`` / ``create-config``."""
