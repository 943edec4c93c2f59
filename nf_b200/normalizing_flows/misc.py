"""Mirror of nisrep/normalizing_flows/misc.py:3-11 (progress bar helper)."""
from tqdm.autonotebook import tqdm


class tqdm_recycled(tqdm):

    def close(self):
        self.reset()

    def really_close(self):
        try:
            self.sp(close=True)
        except (AttributeError, TypeError):
            pass
