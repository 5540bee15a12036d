"""Stand-in for the `progressbar2` package, which the reference imports at pygam/pygam.py:9 and which is not in
this image (no network).  Only ProgressBar()(iterable) is used, by gridsearch (pygam.py:2038-2042).  Nothing of
the reference itself is altered: bench.py --impl reference puts this directory and baseline/_ref (pip install
--target of /root/reference) on sys.path."""


class ProgressBar:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, it):
        return it
