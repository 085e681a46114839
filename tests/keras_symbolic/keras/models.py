from deephar_b200.keras_compat import Model  # noqa: F401
