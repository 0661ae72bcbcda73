import logging as _logging


def get_logger(name):
    return _logging.getLogger("bldfm." + str(name))


def setup_logging(log_dir="logs", **kwargs):
    return None
