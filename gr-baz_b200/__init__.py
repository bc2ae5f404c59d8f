"""gr-baz_b200: B200-native (sm_100a) drop-in for gr-baz's MUSIC direction-of-arrival block.

Only the hot path of /root/reference/lib/baz_music_doa.cc (work()) and its Python helper
(/root/reference/python/music_doa_helper.py) is rebuilt here; see DESIGN.md.
The directory name carries a hyphen, so it is imported through the ``gr_baz_b200`` alias
package at the repo root.
"""
