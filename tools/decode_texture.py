"""Pre-decode a JPEG/PNG texture to the raw .bgr8 format read by the C++ host mirror
(raytrace3_b200/include/raytracer/Element.h): "BGR8 <rows> <cols>\\n" + rows*cols*3 bytes.
cv2.imread is the codec family the reference uses (Element.h:49).
  python tools/decode_texture.py wall.jpg wall.bgr8"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raytrace3_b200.scene import Texture, write_bgr8  # noqa: E402

if __name__ == "__main__":
    write_bgr8(sys.argv[2], Texture(sys.argv[1]))
