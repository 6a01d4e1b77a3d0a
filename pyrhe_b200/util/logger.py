"""Result log: every `_log` line is buffered (and echoed unless suppressed); `_save_log` writes the buffer to
`output_file`.  The buffered text is the result format downstream tools parse, so `_log` keeps the reference's
joining rule (arguments separated by one space, then `end`) -- /root/reference/pyrhe/src/util/logger.py:3-25.
"""
import sys


class Logger:
    def __init__(self, output_file=None, suppress=False, debug_mode=True):
        self.output_file = output_file
        self.suppress = bool(suppress)
        self.debug_mode = bool(debug_mode)
        self.msgs = []

    def _emit(self, text, end="\n"):
        sys.stdout.write(text + end)

    def _debug(self, msg):
        """Diagnostics: printed only in debug mode, never part of the saved log."""
        if self.debug_mode:
            self._emit(str(msg))

    def _log(self, *args, end="\n"):
        line = " ".join(str(a) for a in args)
        self.msgs.append(line + end)
        if not self.suppress:
            self._emit(line, end)

    def _save_log(self):
        if self.output_file is not None:
            with open(self.output_file, "w") as fd:
                fd.write("".join(self.msgs))
