"""oracle/xslt_subset.py against hand-checked XSLT 1.0 / XPath 1.0 behaviour (the recommendations' own rules), on small
stylesheets written here -- so the expander that turns the reference's templates into code is itself checked without the
reference tree -- and, where /root/reference is mounted, against structural facts of the real expansion."""
import os

import pytest

from common import REFERENCE, ROOT  # noqa: F401  (puts the repository on sys.path)
from oracle import xslt_subset as X

NS = 'xmlns:xsl="http://www.w3.org/1999/XSL/Transform" xmlns:m="urn:model"'
DOC = """<?xml version="1.0"?>
<m:model xmlns:m="urn:model" xmlns="urn:plain">
  <m:agent kind="a"><name>wolf</name><size>30</size><m:var><name>x</name><type>float</type></m:var>
     <m:var><name>id</name><type>int</type></m:var><m:var><name>pos</name><type>fvec3</type><len>4</len></m:var></m:agent>
  <m:agent kind="b"><name>sheep</name><size>200</size><m:var><name>y</name><type>double</type></m:var></m:agent>
  <m:agent kind="a"><name>grass</name><size>7</size></m:agent>
</m:model>
"""


def run(tmp_path, body, extra="", doc=DOC, params=None):
    (tmp_path / "s.xslt").write_text(f'<xsl:stylesheet version="1.0" {NS} xmlns:p="urn:plain"><xsl:output method="text"/>'
                                     f'{extra}<xsl:template match="/">{body}</xsl:template></xsl:stylesheet>')
    (tmp_path / "d.xml").write_text(doc)
    return X.Stylesheet(str(tmp_path / "s.xslt")).transform(str(tmp_path / "d.xml"), params)


def test_paths_namespaces_predicates_and_positions(tmp_path):
    out = run(tmp_path, '<xsl:for-each select="m:model/m:agent[@kind=\'a\']"><xsl:value-of select="p:name"/>:'
                        '<xsl:value-of select="position()"/>/<xsl:value-of select="last()"/>'
                        '<xsl:if test="position()!=last()">,</xsl:if></xsl:for-each>')
    assert out == "wolf:1/2,grass:2/2"
    assert run(tmp_path, '<xsl:value-of select="count(//m:var)"/>|<xsl:value-of select="count(m:model/*)"/>|'
                         '<xsl:value-of select="m:model/m:agent[2]/p:name"/>|'
                         '<xsl:value-of select="m:model/m:agent[p:name=\'wolf\']/m:var[not(p:len)][2]/p:name"/>') == "4|3|sheep|id"
    # an unprefixed name test selects the null namespace only; the document's default namespace needs a prefix
    assert run(tmp_path, '<xsl:value-of select="count(m:model/m:agent/name)"/>') == "0"
    # parent steps and variables inside predicates
    assert run(tmp_path, '<xsl:variable name="who" select="\'sheep\'"/><xsl:for-each select="//m:var[p:name=\'x\']">'
                         '<xsl:value-of select="../../m:agent[p:name=$who]/p:size"/></xsl:for-each>') == "200"


def test_existential_comparisons_arithmetic_and_number_formatting(tmp_path):
    # node-set = string is true if ANY node matches; != likewise (so both can hold at once)
    assert run(tmp_path, '<xsl:value-of select="//p:type=\'double\'"/>,<xsl:value-of select="//p:type!=\'double\'"/>,'
                         '<xsl:value-of select="//p:type=\'half\'"/>,<xsl:value-of select="not(//p:nothing)"/>') == "true,true,false,true"
    assert run(tmp_path, '<xsl:value-of select="7 div 2"/>,<xsl:value-of select="6 div 2"/>,<xsl:value-of select="7 mod 4"/>,'
                         '<xsl:value-of select="-7 mod 4"/>,<xsl:value-of select="ceiling(2.1)"/>,<xsl:value-of select="round(2.5)"/>,'
                         '<xsl:value-of select="round(-2.5)"/>,<xsl:value-of select="floor(-0.5)"/>,<xsl:value-of select="1 div 0"/>,'
                         '<xsl:value-of select="\'abc\' * 2"/>') == "3.5,3,3,-3,3,3,-2,-1,Infinity,NaN"
    assert run(tmp_path, '<xsl:value-of select="(//p:size)[1] - 2 * 3"/>,<xsl:value-of select="sum(//p:size)"/>,'
                         '<xsl:value-of select="//p:size &gt; 100"/>,<xsl:value-of select="3 &lt; 2 or 1 = 1.0 and \'a\' != \'b\'"/>') == "24,237,true,true"
    # booleans in arithmetic (the templates' abs helper multiplies by a comparison)
    assert run(tmp_path, '<xsl:variable name="v" select="-4"/><xsl:value-of select="$v * ($v &gt;= 0) - $v * ($v &lt; 0)"/>') == "4"


def test_string_functions(tmp_path):
    assert run(tmp_path, '<xsl:variable name="t" select="\'glm::fvec3\'"/>'
                         '<xsl:value-of select="substring($t, string-length($t) - string-length(\'vec3\') + 1)"/>,'
                         '<xsl:value-of select="substring(\'12345\', 1.5, 2.6)"/>,<xsl:value-of select="substring(\'12345\', 0, 3)"/>,'
                         '<xsl:value-of select="translate(\'1,2,3\', \',\', \'\')"/>,<xsl:value-of select="translate(\'abc\', \'abc\', \'AB\')"/>,'
                         '<xsl:value-of select="contains($t, \'vec\')"/>,<xsl:value-of select="substring-before($t, \'::\')"/>,'
                         '<xsl:value-of select="substring-after($t, \'::\')"/>,<xsl:value-of select="concat(\'a\', 1, true())"/>,'
                         '<xsl:value-of select="normalize-space(\'  a   b \')"/>,<xsl:value-of select="starts-with($t, \'glm\')"/>,'
                         '<xsl:value-of select="local-name(m:model/m:agent)"/>') == "vec3,234,12,123,AB,true,glm,fvec3,a1true,a b,true,agent"
    assert run(tmp_path, '<xsl:value-of select="namespace-uri(m:model)"/>,<xsl:value-of select="name(m:model/m:agent)"/>,'
                         '<xsl:value-of select="namespace-uri(//p:name)"/>,<xsl:value-of select="namespace-uri(//m:agent/@kind)"/>;'
                         '<xsl:value-of select="string(//m:agent[3]/@kind)"/>,<xsl:value-of select="string-length(//p:nothing)"/>') \
        == "urn:model,m:agent,urn:plain,;a,0"


def test_sort_choose_and_whitespace_rules(tmp_path):
    body = ('<xsl:for-each select="//m:agent"><xsl:sort select="p:size" data-type="number" order="descending"/>'
            '<xsl:value-of select="p:name"/><xsl:text> </xsl:text></xsl:for-each>|'
            '<xsl:for-each select="//m:agent"><xsl:sort select="p:size"/><xsl:value-of select="p:size"/>;</xsl:for-each>|'
            '<xsl:for-each select="//m:agent"><xsl:sort select="@kind"/><xsl:sort select="p:name" order="descending"/>'
            '<xsl:value-of select="p:name"/>;</xsl:for-each>')
    assert run(tmp_path, body) == "sheep wolf grass |200;30;7;|wolf;grass;sheep;"
    body = ('<xsl:for-each select="//m:agent">\n   <xsl:choose>\n  <xsl:when test="p:size &gt; 100">big</xsl:when>\n'
            '<xsl:when test="p:size &gt; 10">mid </xsl:when><xsl:otherwise>\n small</xsl:otherwise></xsl:choose></xsl:for-each>')
    # whitespace-only text of the stylesheet is dropped, any other text is kept verbatim including its own whitespace
    assert run(tmp_path, body) == "mid big\n small"


def test_templates_params_scoping_and_recursion(tmp_path):
    extra = ('<xsl:template name="rep"><xsl:param name="n"/><xsl:param name="s" select="\'*\'"/><xsl:param name="unset"/>'
             '<xsl:if test="$n &gt; 0"><xsl:value-of select="$s"/><xsl:call-template name="rep"><xsl:with-param name="n" select="$n - 1"/>'
             '<xsl:with-param name="s" select="$s"/></xsl:call-template></xsl:if>'
             '<xsl:if test="$n = 0 and $unset">never</xsl:if></xsl:template>'
             '<xsl:template match="m:var">[<xsl:value-of select="p:name"/><xsl:apply-templates select="p:len"/>]</xsl:template>'
             '<xsl:template match="m:var/p:len">x<xsl:value-of select="."/></xsl:template>'
             '<xsl:variable name="g" select="count(//m:agent)"/>')
    body = ('<xsl:call-template name="rep"><xsl:with-param name="n" select="$g"/></xsl:call-template>|'
            '<xsl:call-template name="rep"><xsl:with-param name="n" select="2"/><xsl:with-param name="s">ab</xsl:with-param></xsl:call-template>|'
            '<xsl:apply-templates select="//m:agent[1]/m:var"/>|'
            '<xsl:variable name="v" select="1"/><xsl:for-each select="//m:agent"><xsl:variable name="v" select="position() * 10"/>'
            '<xsl:value-of select="$v"/>,</xsl:for-each><xsl:value-of select="$v"/>')
    # a parameter with neither select nor content is the empty string (false); a result-tree fragment is always true
    assert run(tmp_path, body, extra) == "***|abab|[x][id][posx4]|10,20,30,1"
    assert run(tmp_path, '<xsl:variable name="f"><xsl:value-of select="//p:nothing"/></xsl:variable>'
                         '<xsl:variable name="e"/><xsl:value-of select="boolean($f)"/>,<xsl:value-of select="boolean($e)"/>,'
                         '<xsl:value-of select="$f = \'\'"/>') == "true,false,true"


def test_outside_the_subset_fails_loudly(tmp_path):
    with pytest.raises(X.XsltError):
        run(tmp_path, '<xsl:value-of select="document(\'x.xml\')"/>')
    with pytest.raises(X.XsltError):
        run(tmp_path, '<xsl:number value="3"/>')
    with pytest.raises(X.XsltError):
        run(tmp_path, '<b>literal element</b>')
    with pytest.raises(X.XsltError):
        run(tmp_path, '<xsl:value-of select="$undeclared"/>')


@pytest.mark.skipif(not os.path.isdir(f"{REFERENCE}/FLAMEGPU/templates"), reason="the reference tree is not mounted")
def test_reference_templates_expand_completely():
    """All six code templates expand for a spatial, a brute-force and a multi-state model with no template syntax left
    over and with the model's names in the places section 8b of SURVEY.md lists."""
    for name, agent, message in (("CirclesPartitioning_float", "Circle", "location"), ("CirclesBruteForce_float", "Circle", "location"),
                                 ("Keratinocyte", "keratinocyte", "force")):
        xml = f"{REFERENCE}/examples/{name}/src/model/XMLModelFile.xml"
        texts = {t: X.Stylesheet(f"{REFERENCE}/FLAMEGPU/templates/{t}.xslt").transform(xml)
                 for t in ("header", "FLAMEGPU_kernals", "simulation", "io", "main", "functions")}
        for t, text in texts.items():
            assert "<xsl:" not in text and "xmml:" not in text, (name, t)
            assert len(text) > 2000, (name, t)
        h = texts["header"]
        assert f"struct xmachine_memory_{agent}_list" in h and f"#define xmachine_message_{message}_MAX" in h
        assert f"add_{message}_message(xmachine_message_{message}_list*" in h
        assert f"get_first_{message}_message(" in h and "void singleIteration();" in h
        assert texts["simulation"].count("<<<") >= 3 and f"GPUFLAME_" in texts["FLAMEGPU_kernals"]
