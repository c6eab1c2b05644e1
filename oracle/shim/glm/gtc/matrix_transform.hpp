// forwarding header of the glm stand-in (oracle only)
#include <glm/glm.hpp>
