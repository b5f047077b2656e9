"""YAML config loading with the reference's names (project/utils/config_utils.py:5-13,122)."""
import yaml


def load_config(file_path):
    with open(file_path, "r", encoding="utf-8") as f:
        return yaml.safe_load(f)


def save_config(config, file_path):
    with open(file_path, "w", encoding="utf-8") as f:
        yaml.safe_dump(config, f, sort_keys=False)


file_loader = load_config
