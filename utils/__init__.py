"""Host-side utilities of the driver application (mirrors the reference's top-level `utils` package)."""
