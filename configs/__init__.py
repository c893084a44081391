"""Model / dataset configurations: the layer programs the Gram kernels are benchmarked on.
They define the same ``initial_model`` trees and dataset splits as the reference's ``configs/``
(see each module for the file:line), and are selected by name with
``importlib.import_module(f"configs.{name}")`` like the reference scripts do."""
