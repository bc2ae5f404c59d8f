"""The non-code boundary files stay signature-compatible with the reference (run where
/root/reference exists, i.e. in the build container; skipped on the GPU box)."""
import os
import re
import xml.etree.ElementTree as ET

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


def _sig(root):
    return (root.findtext("key"), root.findtext("import"), root.findtext("make"), root.findtext("callback"),
            [(p.findtext("key"), p.findtext("value"), p.findtext("type")) for p in root.findall("param")],
            [(s.findtext("name"), s.findtext("type"), s.findtext("vlen")) for s in root.findall("sink")],
            [(s.findtext("name"), s.findtext("type"), s.findtext("vlen"), s.findtext("optional")) for s in root.findall("source")])


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_grc_descriptor_is_compatible_with_reference():
    a = ET.parse(os.path.join(REF, "grc", "baz_music_doa.xml")).getroot()
    b = ET.parse(os.path.join(ROOT, "grc", "baz_music_doa.xml")).getroot()
    assert _sig(a) == _sig(b)


def _decls(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    return re.sub(r"\s+", " ", text)


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not present")
def test_cpp_surface_matches_reference_header():
    ref = _decls(open(os.path.join(REF, "lib", "baz_music_doa.h")).read())
    ours = _decls(open(os.path.join(ROOT, "lib", "baz_music_doa.h")).read())
    norm = lambda s: re.sub(r"\s*([&*(),])\s*", r"\1", s)
    for must in (
        "class baz_music_doa : public gr::sync_block",
        "typedef boost::shared_ptr<baz_music_doa> baz_music_doa_sptr;",
        "typedef std::vector<gr_complex> antenna_response_t;",
        "typedef std::vector<antenna_response_t> array_response_t;",
        "baz_music_doa_sptr baz_make_music_doa(unsigned int m, unsigned int n, unsigned int nsamples, const array_response_t& array_response, unsigned int resolution);",
        "int work(int noutput_items, gr_vector_const_void_star &input_items, gr_vector_void_star &output_items);",
        "void set_array_response(const array_response_t& array_response);",
    ):
        assert norm(must) in norm(ref), must
        assert norm(must) in norm(ours), must


def test_swig_fragment_keeps_python_name():
    s = open(os.path.join(ROOT, "swig", "baz_music_doa.i")).read()
    assert "GR_SWIG_BLOCK_MAGIC(baz,music_doa)" in s
    assert "std::vector<std::vector<gr_complex> >& array_response" in s
