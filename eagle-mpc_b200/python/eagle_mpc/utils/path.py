"""EAGLE_MPC_YAML_DIR / EAGLE_MPC_ROBOT_DATA_DIR (bindings/python/eagle_mpc/utils/path.py.in): the repo's yaml/ and the
synthetic URDFs under fixtures/urdf, overridable by environment variables of the same names."""
import os

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))))
EAGLE_MPC_YAML_DIR = os.environ.get("EAGLE_MPC_YAML_DIR", os.path.join(_REPO, "yaml"))
EAGLE_MPC_ROBOT_DATA_DIR = os.environ.get("EAGLE_MPC_ROBOT_DATA_DIR", os.path.join(_REPO, "fixtures", "urdf"))
