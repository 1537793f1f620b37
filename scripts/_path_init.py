"""Put the directory that contains ``scripts/`` first on sys.path, so that the ``vision_base`` and
``monodepth`` packages next to this script are the ones the configs resolve (reference: scripts/_path_init.py)."""
import logging
import os
import sys

package_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, package_path)


def manage_package_logging():
    try:
        import coloredlogs
        coloredlogs.install(logging.CRITICAL)
    except ImportError:        # coloredlogs is optional here (absent from the image)
        logging.getLogger().setLevel(logging.CRITICAL)
