"""Progress-bar helper carrying the reference's name (nisrep/normalizing_flows/misc.py): one bar object is
re-used for every epoch of a training run, so ``close()`` — which tqdm's iterator protocol calls at the end
of each loop — only rewinds it, and ``really_close()`` releases it when training is over."""
import tqdm.autonotebook as _autonotebook


class tqdm_recycled(_autonotebook.tqdm):
    """A tqdm bar that survives ``close()``."""

    def close(self):
        self.reset()

    def really_close(self):
        status_printer = getattr(self, "sp", None)      # console bars own a status printer; notebook bars do not
        try:
            if status_printer is not None:
                status_printer(close=True)
            else:
                super().close()
        except (AttributeError, TypeError):
            pass
