from tensorflow.keras.regularizers import L2, l2  # noqa: F401  (model.py:3 imports the module only)
