// examples/main.cpp — the body of the reference's raytracer/main.cpp:19-41, unchanged except
// for the bounded loop counts read from the environment (the reference loops 100000 times) and
// the removed system("pause").  Build: make -C examples.
#include <cstdlib>

#include "Vec3.h"
#include "Obj.h"
#include "Camera.h"
#include "Element.h"
#include "Light.h"
#include "Scene.h"
#include "Raytracer.h"

static int env_int(const char* k, int dflt) { return std::getenv(k) ? std::atoi(std::getenv(k)) : dflt; }

int main()
{
	RayTracer r1;
	Camera camera;
	Vec3 org = Vec3(50, 35, 230);
	camera.init(env_int("SPPM_W", Camera::DEFAULT_WIDTH), env_int("SPPM_H", Camera::DEFAULT_HEIGHT));
	camera.MoveTo(org);
	camera.lookAt(org + Vec3(0, 0.042612, -1), 1 / 1.);
	Scene scene;
	scene.init();

	/*r1.camera = &camera;
	r1.scene = &scene;
	r1.render();*/

	SPPMRayTracer r2;
	r2.sppmtime = env_int("SPPM_ITERS", 1);
	r2.cnt = env_int("SPPM_CNT", 10);
	r2.onetime = env_int("SPPM_ONETIME", 100000);
	if (std::getenv("SPPM_JITTER")) r2.jitter = std::atof(std::getenv("SPPM_JITTER"));
	r2.shared_statistics = env_int("SPPM_SHARED", 0) != 0;  // SPPM proper instead of the average of PPM renders
	r2.render(scene, camera);

	double sum = 0;
	for (double v : r2.film) sum += v;  // render() takes the camera by value (as upstream): read the renderer's film
	std::cout.precision(17);
	std::cout << "film sum " << sum << std::endl;
	return 0;
}
