// xml_dom.h — a small in-memory XML element tree, enough for the diffusion-curve scene files.
// Fills the role rapidxml plays in the reference (optixHello.cpp:108-111, doc.parse<0>): elements and
// attributes only; DOCTYPE, comments, processing instructions and character data are skipped, exactly
// the nodes parse<0> does not create. Children and attributes keep document order; lookup by name is
// case-sensitive and returns the first match, like first_node(name)/first_attribute(name).
#ifndef RDC_XML_DOM_H
#define RDC_XML_DOM_H

#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rdc {

struct XmlError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct XmlElement {
  std::string name;
  std::vector<std::pair<std::string, std::string>> attrs;
  std::vector<std::unique_ptr<XmlElement>> children;

  const std::string* attr(const char* key) const {
    for (auto& a : attrs)
      if (a.first == key) return &a.second;
    return nullptr;
  }
  const XmlElement* child(const char* key) const {
    for (auto& c : children)
      if (c->name == key) return c.get();
    return nullptr;
  }
};

// Parses a whole document held in memory and returns its first top-level element.
std::unique_ptr<XmlElement> xml_parse(const char* text, size_t len);
// Reads the file and parses it. Throws XmlError when the file cannot be read or is malformed.
std::unique_ptr<XmlElement> xml_parse_file(const std::string& path);
// Canonical dump (one line per element, attributes in document order) used by the parser tests.
void xml_dump(const XmlElement& e, int depth, std::string& out);

}  // namespace rdc

#endif
