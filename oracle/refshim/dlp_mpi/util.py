def ensure_single_thread_numeric():
    pass
