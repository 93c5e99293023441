"""Drop-in shim: `python -m core.training --config <yaml>` resolves to the B200 hot path."""
