"""Message buffer -> stdout + log file (interface of /root/reference/pyrhe/src/util/logger.py:3-25).

The buffered text IS the result format the reference's tests parse, so `_log`
joins its arguments with one space and appends `end`, exactly like the reference.
"""


class Logger:
    def __init__(self, output_file=None, suppress=False, debug_mode=True):
        self.msgs = []
        self.output_file = output_file
        self.suppress = suppress
        self.debug_mode = debug_mode

    def _debug(self, msg):
        if self.debug_mode:
            print(msg)

    def _log(self, *args, end="\n"):
        text = " ".join(map(str, args))
        self.msgs.append(text + end)
        if not self.suppress:
            print(text, end=end)

    def _save_log(self):
        if self.output_file is None:
            return
        with open(self.output_file, "w") as fd:
            fd.writelines(self.msgs)
