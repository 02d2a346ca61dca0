def __getattr__(name):
    raise RuntimeError("matplotlib stub: pyplot.%s is not available" % name)
