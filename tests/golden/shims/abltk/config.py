from pathlib import Path

DATA_DIR = CACHE_DIR = CONFIG_DIR = OUTPUT_DIR = LOG_DIR = None


def init(root):
    global DATA_DIR, CACHE_DIR, CONFIG_DIR, OUTPUT_DIR, LOG_DIR
    root = Path(root)
    DATA_DIR = str(root / "data")
    CACHE_DIR = str(root / ".cache")
    CONFIG_DIR = str(root / "config")
    OUTPUT_DIR = str(root / "output")
    LOG_DIR = str(root / "logs")
