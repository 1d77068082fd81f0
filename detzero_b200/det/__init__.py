from .centerpoint import CenterPoint, build_network, cp_modules, load_data_to_gpu  # noqa: F401
