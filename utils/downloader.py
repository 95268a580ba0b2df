"""Download a file with a progress line; a file that already exists is left alone."""
import os
from urllib.error import URLError
from urllib.request import urlretrieve


def show_progress(blk_num, blk_sz, tot_sz):
    print("Progress: %.1f %%" % (100. * blk_num * blk_sz / tot_sz), end="\r", flush=True)


def download_url(url, file_path):
    folder = os.path.dirname(file_path)
    if folder and not os.path.exists(folder):
        os.makedirs(folder)
    try:
        if os.path.exists(file_path):
            print("{} already exists.".format(file_path))
            return
        print("Downloading {} to {}".format(url, file_path))
        try:
            urlretrieve(url, file_path, show_progress)
        except URLError:
            raise RuntimeError("Error downloading resource!")
        finally:
            print()
    except KeyboardInterrupt:
        print("Interrupted")
