"""Reference import path: `from model_mip import ...` (barf/model_mip.py); the classes live in
model_camera_calibration.py."""
from .model_camera_calibration import BarfModel, CameraCalibrationModel, MipBarf, MipNeRF  # noqa: F401
