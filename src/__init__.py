"""Import-compatibility package: `from src.lib.X import ...` resolves to the B200-backed mirror of the reference's src/lib."""
