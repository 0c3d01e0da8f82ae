"""pytest-mock's ``mocker`` fixture, as far as the reference's tests use it (tests/test_fusion.py:371-420, 590-640), on unittest.mock --
pytest-mock is not installed in this image.  Copied as conftest.py next to the reference's test files by tests/test_reference_suite_cpu.py."""
from unittest import mock

import pytest


class _Mocker:
    Mock = mock.Mock
    MagicMock = mock.MagicMock

    def __init__(self):
        self._patches = []

    def patch(self, target, *args, **kwargs):
        p = mock.patch(target, *args, **kwargs)
        m = p.start()
        self._patches.append(p)
        return m

    def stopall(self):
        for p in reversed(self._patches):
            p.stop()


@pytest.fixture
def mocker():
    m = _Mocker()
    yield m
    m.stopall()
