// DriverIo.hpp — the input and output formats of the reference's driver (src/OpenMps/Main.cpp) without Boost:
//   * the XML run description: <openmps><condition>, <environment> (numbers in `value` attributes) and
//     <particles type="csv"> with the particle table as element text        (Main.cpp:186-274, Boost.PropertyTree there)
//   * the embedded particle CSV: header-name driven columns, blanks and tabs ignored, empty lines skipped (Main.cpp:70-183)
//   * the result CSV  result/particles_%05d.csv: "Type, x, z, u, w, p, n" and default `ostream << double` text (Main.cpp:31-67)
//   * the progress line                                                                                      (Main.cpp:334-353)
// Not part of the reference's header set: SURVEY.md §8f rank 1 ("next" row).  Same messages for the same malformed inputs.
#ifndef OPENMPS_DRIVER_IO_INCLUDED
#define OPENMPS_DRIVER_IO_INCLUDED

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "defines.hpp"
#include "ComputingCondition.hpp"
#include "Environment.hpp"
#include "Particle.hpp"

namespace { namespace OpenMps { namespace DriverIo
{
	// ---- a minimal XML reader: elements, attributes, text, comments, declarations, CDATA, the five predefined entities -------
	struct XmlNode
	{
		std::string name;
		std::vector<std::pair<std::string, std::string>> attributes;
		std::string text;                               // concatenated character data of this element (children excluded)
		std::vector<std::unique_ptr<XmlNode>> children;

		const XmlNode* Child(const std::string& n) const
		{
			for (const auto& c : children) if (c->name == n) return c.get();
			return nullptr;
		}
		const std::string* Attribute(const std::string& n) const
		{
			for (const auto& a : attributes) if (a.first == n) return &a.second;
			return nullptr;
		}
	};

	class XmlReader final
	{
		const std::string& s;
		std::size_t i = 0;

		[[noreturn]] void Fail(const char* what) const
		{
			std::size_t line = 1;
			for (std::size_t k = 0; k < i && k < s.size(); k++) if (s[k] == '\n') line++;
			throw std::runtime_error(std::string("XML parse error (line ") + std::to_string(line) + "): " + what);
		}
		bool StartsWith(const char* t) const { return s.compare(i, std::strlen(t), t) == 0; }
		void SkipSpace() { while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++; }
		void SkipUntil(const char* t)
		{
			const auto k = s.find(t, i);
			if (k == std::string::npos) Fail("unterminated construct");
			i = k + std::strlen(t);
		}
		static bool NameChar(const char c) { return std::isalnum(static_cast<unsigned char>(c)) || c == '_' || c == '-' || c == '.' || c == ':'; }
		std::string Name()
		{
			const auto b = i;
			while (i < s.size() && NameChar(s[i])) i++;
			if (i == b) Fail("name expected");
			return s.substr(b, i - b);
		}
		static void AppendDecoded(std::string& out, const std::string& raw)
		{
			for (std::size_t k = 0; k < raw.size(); k++)
			{
				if (raw[k] != '&') { out.push_back(raw[k]); continue; }
				const auto e = raw.find(';', k);
				const std::string ent = (e == std::string::npos) ? std::string() : raw.substr(k + 1, e - k - 1);
				if (ent == "lt") out.push_back('<');
				else if (ent == "gt") out.push_back('>');
				else if (ent == "amp") out.push_back('&');
				else if (ent == "quot") out.push_back('"');
				else if (ent == "apos") out.push_back('\'');
				else if (ent.size() > 1 && ent[0] == '#')
				{
					const long code = (ent[1] == 'x') ? std::strtol(ent.c_str() + 2, nullptr, 16) : std::strtol(ent.c_str() + 1, nullptr, 10);
					if (code < 0x80) out.push_back(static_cast<char>(code)); else out.push_back('?');
				}
				else { out.push_back('&'); continue; }
				k = e;
			}
		}
		// skips comments, processing instructions and DOCTYPE; returns false at a real tag or at text
		bool SkipMisc()
		{
			if (StartsWith("<!--")) { i += 4; SkipUntil("-->"); return true; }
			if (StartsWith("<?")) { i += 2; SkipUntil("?>"); return true; }
			if (StartsWith("<!DOCTYPE")) { SkipUntil(">"); return true; }
			return false;
		}
		std::unique_ptr<XmlNode> Element()
		{
			// s[i] == '<' and a name follows
			i++;
			auto node = std::make_unique<XmlNode>();
			node->name = Name();
			for (;;)
			{
				SkipSpace();
				if (i >= s.size()) Fail("unterminated start tag");
				if (StartsWith("/>")) { i += 2; return node; }
				if (s[i] == '>') { i++; break; }
				std::string key = Name();
				SkipSpace();
				if (i >= s.size() || s[i] != '=') Fail("'=' expected after attribute name");
				i++;
				SkipSpace();
				if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) Fail("quoted attribute value expected");
				const char q = s[i++];
				const auto e = s.find(q, i);
				if (e == std::string::npos) Fail("unterminated attribute value");
				std::string value;
				AppendDecoded(value, s.substr(i, e - i));
				i = e + 1;
				node->attributes.emplace_back(std::move(key), std::move(value));
			}
			// content
			for (;;)
			{
				if (i >= s.size()) Fail("unterminated element");
				if (s[i] != '<')
				{
					const auto e = s.find('<', i);
					if (e == std::string::npos) Fail("unterminated element");
					AppendDecoded(node->text, s.substr(i, e - i));
					i = e;
					continue;
				}
				if (SkipMisc()) continue;
				if (StartsWith("<![CDATA["))
				{
					i += 9;
					const auto e = s.find("]]>", i);
					if (e == std::string::npos) Fail("unterminated CDATA");
					node->text.append(s, i, e - i);
					i = e + 3;
					continue;
				}
				if (StartsWith("</"))
				{
					i += 2;
					const std::string close = Name();
					if (close != node->name) Fail("mismatched end tag");
					SkipSpace();
					if (i >= s.size() || s[i] != '>') Fail("'>' expected");
					i++;
					return node;
				}
				node->children.push_back(Element());
			}
		}

	public:
		explicit XmlReader(const std::string& text) : s(text) {}

		std::unique_ptr<XmlNode> Document()
		{
			if (s.compare(0, 3, "\xEF\xBB\xBF") == 0) i = 3; // UTF-8 byte order mark
			for (;;)
			{
				SkipSpace();
				if (i >= s.size()) Fail("no root element");
				if (s[i] != '<') Fail("text outside the root element");
				if (SkipMisc()) continue;
				return Element();
			}
		}
	};

	inline std::unique_ptr<XmlNode> ReadXmlFile(const std::string& filename)
	{
		std::ifstream in(filename, std::ios::binary);
		if (!in) throw std::runtime_error(filename + ": cannot open file"); // property_tree: xml_parser_error "cannot open file"
		std::stringstream buffer;
		buffer << in.rdbuf();
		const std::string text = buffer.str();
		return XmlReader(text).Document();
	}

	// "openmps.environment.l_0" -> node ; throws like ptree_bad_path when it does not exist
	inline const XmlNode& Path(const XmlNode& root, const std::string& path)
	{
		const XmlNode* cur = nullptr;
		std::size_t b = 0;
		bool first = true;
		while (b <= path.size())
		{
			const auto e = path.find('.', b);
			const std::string part = path.substr(b, (e == std::string::npos) ? std::string::npos : e - b);
			if (first) { if (root.name != part) throw std::runtime_error("No such node (" + path + ")"); cur = &root; first = false; }
			else
			{
				cur = cur->Child(part);
				if (!cur) throw std::runtime_error("No such node (" + path + ")");
			}
			if (e == std::string::npos) break;
			b = e + 1;
		}
		return *cur;
	}

	inline const std::string& ValueAttribute(const XmlNode& root, const std::string& path)
	{
		const auto* v = Path(root, path).Attribute("value");
		if (!v) throw std::runtime_error("No such node (" + path + ".<xmlattr>.value)");
		return *v;
	}

	// like property_tree's stream translator: leading white space allowed, the WHOLE remaining text must be the number
	inline double GetDouble(const XmlNode& root, const std::string& path)
	{
		const std::string& v = ValueAttribute(root, path);
		const char* b = v.c_str();
		char* e = nullptr;
		errno = 0;
		const double x = std::strtod(b, &e);
		if (e == b) throw std::runtime_error("conversion of data to type \"double\" failed (" + path + ")");
		while (*e == ' ' || *e == '\t' || *e == '\n' || *e == '\r') e++;
		if (*e != '\0') throw std::runtime_error("conversion of data to type \"double\" failed (" + path + ")");
		return x;
	}
	inline std::size_t GetSize(const XmlNode& root, const std::string& path)
	{
		const std::string& v = ValueAttribute(root, path);
		const char* b = v.c_str();
		while (*b == ' ' || *b == '\t' || *b == '\n' || *b == '\r') b++;
		char* e = nullptr;
		if (*b == '-' ) throw std::runtime_error("conversion of data to type \"unsigned long\" failed (" + path + ")");
		const unsigned long long x = std::strtoull(b, &e, 10);
		if (e == b) throw std::runtime_error("conversion of data to type \"unsigned long\" failed (" + path + ")");
		while (*e == ' ' || *e == '\t' || *e == '\n' || *e == '\r') e++;
		if (*e != '\0') throw std::runtime_error("conversion of data to type \"unsigned long\" failed (" + path + ")");
		return static_cast<std::size_t>(x);
	}

	// ---- Main.cpp:254-274 ----------------------------------------------------------------------------------------------------
	inline ComputingCondition LoadCondition(const XmlNode& xml)
	{
		const auto startTime = GetDouble(xml, "openmps.condition.startTime");
		const auto endTime = GetDouble(xml, "openmps.condition.endTime");
		const auto outputInterval = GetDouble(xml, "openmps.condition.outputInterval");
		const auto eps = GetDouble(xml, "openmps.condition.eps");
		return ComputingCondition(eps, startTime, endTime, outputInterval);
	}

	// ---- Main.cpp:202-251 (default build: MPS_SPP defined, so surfaceRatio is not read; no c, no tooNear*) ---------------------
	inline Environment LoadEnvironment(const XmlNode& xml, const double outputInterval)
	{
		const auto l_0 = GetDouble(xml, "openmps.environment.l_0");
		const auto minStepCountPerOutput = GetSize(xml, "openmps.environment.minStepCountPerOutput");
		const double courant = GetDouble(xml, "openmps.environment.courant");
		const double g = GetDouble(xml, "openmps.environment.g");
		const double rho = GetDouble(xml, "openmps.environment.rho");
		const double nu = GetDouble(xml, "openmps.environment.nu");
		const double r_eByl_0 = GetDouble(xml, "openmps.environment.r_eByl_0");
		const double minX = GetDouble(xml, "openmps.environment.minX");
#ifdef DIM3
		const double minY = GetDouble(xml, "openmps.environment.minY");
#endif
		const double minZ = GetDouble(xml, "openmps.environment.minZ");
		const double maxX = GetDouble(xml, "openmps.environment.maxX");
#ifdef DIM3
		const double maxY = GetDouble(xml, "openmps.environment.maxY");
#endif
		const double maxZ = GetDouble(xml, "openmps.environment.maxZ");
		return Environment(outputInterval / static_cast<double>(minStepCountPerOutput), courant, g, rho, nu, r_eByl_0, l_0,
#ifdef DIM3
			minX, minY, minZ, maxX, maxY, maxZ
#else
			minX, minZ, maxX, maxZ
#endif
		);
	}

	// ---- Main.cpp:70-183 -----------------------------------------------------------------------------------------------------
	inline std::vector<Particle> InputFromCsv(const std::string& csv, std::ostream& log = std::cout)
	{
		std::vector<Particle> particles;

		// lines (split at '\n' only, like the reference; a trailing '\r' stays part of the last item)
		std::vector<std::pair<std::size_t, std::size_t>> lines; // begin, length
		for (std::size_t b = 0;;)
		{
			const auto e = csv.find('\n', b);
			if (e == std::string::npos) { lines.emplace_back(b, csv.size() - b); break; }
			lines.emplace_back(b, e - b);
			b = e + 1;
		}
		// a line with its blanks and tabs removed, cut at the commas (an empty line gives no items; "a,,b" keeps the empty item)
		const auto GetItems = [&csv](const std::pair<std::size_t, std::size_t>& line)
		{
			std::string str;
			str.reserve(line.second);
			for (std::size_t k = 0; k < line.second; k++)
			{
				const char c = csv[line.first + k];
				if (c != ' ' && c != '\t') str.push_back(c);
			}
			std::vector<std::string> data;
			if (!str.empty())
			{
				std::size_t b = 0;
				for (;;)
				{
					const auto e = str.find(',', b);
					data.push_back(str.substr(b, (e == std::string::npos) ? std::string::npos : e - b));
					if (e == std::string::npos) break;
					b = e + 1;
				}
			}
			return data;
		};

		// leading empty lines are skipped (empty = zero characters, Main.cpp:78-80)
		std::size_t li = 0;
		while (li < lines.size() && lines[li].second == 0) li++;
		if (li == lines.size()) throw std::runtime_error("Some header item doesn't exist");

		constexpr std::size_t HEADER_NOT_FOUND = 0;
		std::unordered_map<std::string, std::size_t> header(
		{
			{ "Type", HEADER_NOT_FOUND }, { "x", HEADER_NOT_FOUND },
#ifdef DIM3
			{ "y", HEADER_NOT_FOUND },
#endif
			{ "z", HEADER_NOT_FOUND }, { "u", HEADER_NOT_FOUND },
#ifdef DIM3
			{ "v", HEADER_NOT_FOUND },
#endif
			{ "w", HEADER_NOT_FOUND }, { "p", HEADER_NOT_FOUND }, { "n", HEADER_NOT_FOUND },
		});
		{
			const auto headerItems = GetItems(lines[li++]);
			for (std::size_t i = 0; i < headerItems.size(); i++)
			{
				const auto it = header.find(headerItems[i]);
				if (it == header.end()) throw std::runtime_error("Illegal header item in input csv");
				it->second = i + 1; // column + 1: 0 means "absent"
			}
			for (auto& item : header)
			{
				if (item.second == HEADER_NOT_FOUND) throw std::runtime_error("Some header item doesn't exist");
				item.second--;
			}
		}
		const std::size_t cType = header["Type"], cX = header["x"], cZ = header["z"], cU = header["u"], cW = header["w"], cP = header["p"], cN = header["n"];
#ifdef DIM3
		const std::size_t cY = header["y"], cV = header["v"];
#endif
		particles.reserve(lines.size() - li);
		for (; li < lines.size(); li++)
		{
			const auto data = GetItems(lines[li]);
			if (data.empty()) continue; // empty lines are skipped
			// std::vector::operator[] out of range is undefined in the reference; here a short row is an error
			const auto item = [&data](const std::size_t c) -> const std::string&
			{
				if (c >= data.size()) throw std::runtime_error("Too few items in a row of input csv");
				return data[c];
			};
			Particle particle(static_cast<Particle::Type>(std::stoi(item(cType))));
			particle.X()[AXIS_X] = std::stod(item(cX));
#ifdef DIM3
			particle.X()[AXIS_Y] = std::stod(item(cY));
#endif
			particle.X()[AXIS_Z] = std::stod(item(cZ));
			particle.U()[AXIS_X] = std::stod(item(cU));
#ifdef DIM3
			particle.U()[AXIS_Y] = std::stod(item(cV));
#endif
			particle.U()[AXIS_Z] = std::stod(item(cW));
			particle.P() = std::stod(item(cP));
			particle.N() = std::stod(item(cN));
			particles.push_back(std::move(particle));
		}
		log << particles.size() << " particles" << std::endl;
		return particles;
	}

	// ---- Main.cpp:186-199 ----------------------------------------------------------------------------------------------------
	inline std::vector<Particle> LoadParticles(const XmlNode& xml, std::ostream& log = std::cout)
	{
		const auto& node = Path(xml, "openmps.particles");
		const auto* type = node.Attribute("type");
		if (type && *type == "csv") return InputFromCsv(node.text, log);
		throw std::runtime_error("Not Implemented!");
	}

	// ---- Main.cpp:31-67: same bytes as `output << double` with the default format (precision 6, %g) ---------------------------
	inline const char* CsvHeader()
	{
#ifdef DIM3
		return "Type, x, y, z, u, v, w, p, n";
#else
		return "Type, x, z, u, w, p, n";
#endif
	}

	// appends one row; returns false for a Disabled particle (the caller counts the others)
	inline bool AppendCsvRow(std::string& out, const Particle& particle)
	{
		char buf[32 + 16 * 32];
		int k = std::snprintf(buf, sizeof(buf), "%d, ", static_cast<int>(particle.TYPE()));
		for (std::size_t d = 0; d < DIM; d++) k += std::snprintf(buf + k, sizeof(buf) - k, "%g, ", particle.X()[d]);
		for (std::size_t d = 0; d < DIM; d++) k += std::snprintf(buf + k, sizeof(buf) - k, "%g, ", particle.U()[d]);
		k += std::snprintf(buf + k, sizeof(buf) - k, "%g, %g\n", particle.P(), particle.N());
		out.append(buf, static_cast<std::size_t>(k));
		return particle.TYPE() != Particle::Type::Disabled;
	}

	inline std::string CsvFileName(const std::size_t outputCount, const std::string& directory = "result")
	{
		char name[64];
		std::snprintf(name, sizeof(name), "/particles_%05zu.csv", outputCount);
		return directory + name;
	}

	// writes the table; returns the number of particles that are not Disabled
	inline std::size_t OutputToCsv(const std::vector<Particle>& particles, const std::size_t outputCount, const std::string& directory = "result")
	{
		std::string text;
		text.reserve(particles.size() * (DIM == 3 ? 100 : 72) + 64);
		text.append(CsvHeader());
		text.push_back('\n');
		std::size_t nonDisableCount = 0;
		for (const auto& particle : particles) nonDisableCount += AppendCsvRow(text, particle) ? 1 : 0;
		std::ofstream output(CsvFileName(outputCount, directory), std::ios::binary);
		output.write(text.data(), static_cast<std::streamsize>(text.size()));
		return nonDisableCount;
	}

	// ---- Main.cpp:334-353: "#%3$05d: t=%1$8.4lf (%2$05d), %10$12d particles, @ %4$02d/%5$02d %6$02d:%7$02d:%8$02d (%9$8.2lf)" -----
	inline std::string ProgressLine(const double tComputer, const std::size_t iteration, const std::size_t outputCount, const std::size_t count,
		const std::tm& now, const double elapsedSeconds)
	{
		char line[192];
		std::snprintf(line, sizeof(line), "#%05zu: t=%8.4lf (%05zu), %12zu particles, @ %02d/%02d %02d:%02d:%02d (%8.2lf)",
			outputCount, tComputer, iteration, count, now.tm_mon + 1, now.tm_mday, now.tm_hour, now.tm_min, now.tm_sec, elapsedSeconds);
		return line;
	}
}}}
#endif
