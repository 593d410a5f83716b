// ref_xml_dump.cpp — TEST INFRASTRUCTURE: dumps the element tree that the reference's XML library
// (rapidxml 1.13, vendored under /root/reference/support/rapidxml, used at optixHello.cpp:108-111 with
// doc.parse<0>) builds for a file, in the same canonical text form as the product's rdc_xml_dump_file().
// Built into oracle/_ref/ from the reference's headers where they lie; only tests run it.
#include <cstdio>
#include <string>

#include <rapidxml/rapidxml.hpp>
#include <rapidxml/rapidxml_utils.hpp>

static void dump(rapidxml::xml_node<>* n, int depth, std::string& out) {
  out.append((size_t)depth, ' ');
  out.append(n->name(), n->name_size());
  for (auto* a = n->first_attribute(); a; a = a->next_attribute()) {
    out += ' ';
    out.append(a->name(), a->name_size());
    out += '=';
    out.append(a->value(), a->value_size());
  }
  out += '\n';
  for (auto* c = n->first_node(); c; c = c->next_sibling())
    if (c->type() == rapidxml::node_element) dump(c, depth + 1, out);
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  try {
    rapidxml::file<> file(argv[1]);
    rapidxml::xml_document<> doc;
    doc.parse<0>(file.data());
    std::string out;
    dump(doc.first_node(), 0, out);
    std::fwrite(out.data(), 1, out.size(), stdout);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return 2;
  }
  return 0;
}
