def enable():
    pass


def disable():
    pass


def set_keepalive_time(t):
    pass
